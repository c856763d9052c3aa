# end-of-session record: full GPU suite, default bench line (with extras and the CPU arm), launch list, smoke
T=${1:-r03e}
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench.json')); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['roofline']['step_share_ms'], 'frac', round(d['roofline']['frac'],3), round(d['roofline']['issued_frac'],3)); print(json.dumps(d.get('extra'), indent=0)[:1500]); print(d.get('cpu_baseline'))"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2

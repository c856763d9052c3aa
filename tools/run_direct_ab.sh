# relaxed cluster waits + the peer's halo loads completing on the leader's barrier (default) against the relay thread
# (DAI_TC_RELAY=1): layer tests, parity suite, counters, interleaved bench runs, then one full default bench line
set -x
T=r02m
timeout 300 python -m pytest tests/test_gpu_layers.py -m gpu -q 2>&1 | tail -4
timeout 900 python -m pytest tests -m "gpu and not fullsize" -q > gpurun_out/${T}_pytest.log 2>&1; tail -5 gpurun_out/${T}_pytest.log
DAI_TC_RELAY=1 timeout 300 python -m pytest tests/test_gpu_layers.py -m gpu -q 2>&1 | tail -2
DAI_GRAPHS=0 DAI_TC_COUNTERS=1 timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 2 --quick > /dev/null 2> gpurun_out/${T}_counters_direct.log
DAI_TC_RELAY=1 DAI_GRAPHS=0 DAI_TC_COUNTERS=1 timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 2 --quick > /dev/null 2> gpurun_out/${T}_counters_relay.log
grep "tc counters" gpurun_out/${T}_counters_direct.log | grep "pair" | sort | uniq -c | sort -rn | head -4 | cut -c1-420
grep "tc counters" gpurun_out/${T}_counters_relay.log | grep "ct2+ct3" | head -1 | cut -c1-420
for rep in 1 2; do
  timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/${T}_direct_$rep.json 2>/dev/null
  DAI_TC_RELAY=1 timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/${T}_relay_$rep.json 2>/dev/null
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02m_*_?.json')):
    try:
        d = json.load(open(f))
        print(f, round(d['value'], 1), {k: round(v, 2) for k, v in d['roofline'].get('step_share_ms', {}).items()})
    except Exception as e:
        print(f, 'ERR', e)
PY
timeout 900 python bench.py > gpurun_out/${T}_bench_full.json 2> gpurun_out/${T}_bench_full.err; tail -c 600 gpurun_out/${T}_bench_full.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02m_bench_full.json'))
print(round(d['value'], 1), round(d['e2e']['value'], 1), {k: (round(v['value'], 2) if 'value' in v else v) for k, v in d.get('extra', {}).items()})
PY

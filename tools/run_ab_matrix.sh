# tests, then an A/B matrix of the switches (each a separate process; interleaved twice to see run-to-run noise)
set -x
timeout 300 python -m pytest tests/test_gpu_layers.py -m gpu -q 2>&1 | tail -5
timeout 900 python -m pytest tests -m "gpu and not fullsize" -q > gpurun_out/r02g_pytest.log 2>&1; tail -15 gpurun_out/r02g_pytest.log
for rep in 1 2; do
  timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r02g_default_$rep.json 2> gpurun_out/r02g_default_$rep.err
  DAI_GRAPHS=0 timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r02g_nograph_$rep.json 2>/dev/null
  DAI_TC_FUSE23=0 timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r02g_unfused_$rep.json 2>/dev/null
  DAI_TC_L2PERSIST=1 timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r02g_l2persist_$rep.json 2> gpurun_out/r02g_l2persist_$rep.err
done
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02g_extras.json 2> gpurun_out/r02g_extras.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02g_*.json')):
    try:
        d = json.load(open(f))
        print(f, round(d['value'], 1), 'launches', d['gpu_launches'], d['roofline'].get('step_share_ms'), json.dumps(d.get('extra', {}))[:600])
    except Exception as e:
        print(f, 'ERR', e)
PY
grep "L2 persistence" gpurun_out/r02g_l2persist_1.err

# tests, then an A/B matrix of the switches (each a separate process; interleaved twice to see run-to-run noise)
set -x
timeout 300 python -m pytest tests/test_gpu_layers.py -m gpu -q 2>&1 | tail -5
timeout 900 python -m pytest tests -m "gpu and not fullsize" -q > gpurun_out/r02g_pytest.log 2>&1; tail -6 gpurun_out/r02g_pytest.log
for rep in 1 2; do
  timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r02g_default_$rep.json 2> gpurun_out/r02g_default_$rep.err
  DAI_GRAPHS=0 timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r02g_nograph_$rep.json 2>/dev/null
  DAI_TC_FUSE23=1 timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r02g_fused_$rep.json 2>/dev/null
done
DAI_GRAPHS=0 DAI_TC_COUNTERS=1 DAI_TC_FUSE23=1 timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 2 --quick > /dev/null 2> gpurun_out/r02g_counters_fused.log
DAI_GRAPHS=0 DAI_TC_COUNTERS=1 timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 2 --quick > /dev/null 2> gpurun_out/r02g_counters_sep.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02g_*.json')):
    try:
        d = json.load(open(f))
        print(f, round(d['value'], 1), 'launches', d['gpu_launches'], {k: round(v, 2) for k, v in d['roofline'].get('step_share_ms', {}).items()})
    except Exception as e:
        print(f, 'ERR', e)
PY
grep "tc counters" gpurun_out/r02g_counters_fused.log | grep pair | head -2
grep "tc counters" gpurun_out/r02g_counters_sep.log | head -8

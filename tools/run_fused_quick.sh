set -x
timeout 300 python -m pytest tests/test_gpu_layers.py -m gpu -q 2>&1 | tail -4
timeout 600 python -m pytest tests -m "gpu and not fullsize" -q -x 2>&1 | tail -4
DAI_GRAPHS=0 DAI_TC_COUNTERS=1 timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 2 --quick > /dev/null 2> gpurun_out/r02i_counters_fused.log
grep "tc counters" gpurun_out/r02i_counters_fused.log | grep pair | head -2
for rep in 1 2; do
  timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r02i_default_$rep.json 2>/dev/null
  DAI_TC_FC4_PAIR=0 timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r02i_nofc4pair_$rep.json 2>/dev/null
  DAI_TC_FUSE23=0 DAI_TC_FC4_PAIR=0 timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r02i_r1kernels_$rep.json 2>/dev/null
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02i_*.json')):
    d = json.load(open(f))
    print(f, round(d['value'], 1), {k: round(v, 2) for k, v in d['roofline'].get('step_share_ms', {}).items()})
PY

# parity of the step path + bench at R=16 and R=1 (tag = $1)
T=${1:-r02u}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_layers.py -m gpu -q -x 2>&1 | tail -3
for rep in 1 2; do
  timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/${T}_bench_$rep.json 2>gpurun_out/${T}_bench_$rep.err
done
timeout 300 python bench.py --roots 1 --no-extras --no-cpu-baseline > gpurun_out/${T}_r1.json 2>/dev/null
timeout 300 python bench.py --roots 1 --samples 100 --horizon 15 --no-extras --no-cpu-baseline > gpurun_out/${T}_c5rank.json 2>/dev/null
python - $T <<'PY'
import json, glob, sys
for f in sorted(glob.glob('gpurun_out/%s_*.json' % sys.argv[1])):
    try:
        d = json.load(open(f))
        print(f, round(d['value'], 1), round(d['ms_per_step'], 3), {k: round(v, 2) for k, v in d['roofline'].get('step_share_ms', {}).items()})
    except Exception as e:
        print(f, 'ERR', e)
PY

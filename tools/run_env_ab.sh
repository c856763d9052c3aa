# interleaved A/B of one environment switch (VAR=VALUE is the B arm) after the GPU parity tests: bash tools/run_env_ab.sh TAG VAR=VALUE [full]
T=$1; KV=$2
if [ "$3" = "full" ]; then timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3; else timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_layers.py -m gpu -q -x 2>&1 | tail -3; fi
for rep in 1 2; do
  timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/${T}_A_R16_$rep.json 2>gpurun_out/${T}_A_R16_$rep.err
  env $KV timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/${T}_B_R16_$rep.json 2>gpurun_out/${T}_B_R16_$rep.err
  timeout 300 python bench.py --roots 1 --no-extras --no-cpu-baseline > gpurun_out/${T}_A_R1_$rep.json 2>/dev/null
  env $KV timeout 300 python bench.py --roots 1 --no-extras --no-cpu-baseline > gpurun_out/${T}_B_R1_$rep.json 2>/dev/null
done
python - $T <<'PY'
import json, glob, sys
for f in sorted(glob.glob('gpurun_out/%s_*.json' % sys.argv[1])):
    try:
        d = json.load(open(f))
        print(f, round(d['value'], 1), round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1), {k: round(v, 2) for k, v in d['roofline'].get('step_share_ms', {}).items()})
    except Exception as e:
        print(f, 'ERR', e)
PY

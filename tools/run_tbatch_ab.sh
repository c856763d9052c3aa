# time-batched horizon on/off: full GPU test suite with it on, then interleaved benches at R=16, R=1 and one rank's share of configs[4]
T=${1:-r02x}
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
for rep in 1 2; do
  for tb in 1 0; do
    DAI_TBATCH=$tb timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/${T}_tb${tb}_R16_$rep.json 2>gpurun_out/${T}_tb${tb}_R16_$rep.err
    DAI_TBATCH=$tb timeout 300 python bench.py --roots 1 --no-extras --no-cpu-baseline > gpurun_out/${T}_tb${tb}_R1_$rep.json 2>gpurun_out/${T}_tb${tb}_R1_$rep.err
    DAI_TBATCH=$tb timeout 300 python bench.py --roots 1 --samples 100 --horizon 15 --no-extras --no-cpu-baseline > gpurun_out/${T}_tb${tb}_c5rank_$rep.json 2>/dev/null
  done
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

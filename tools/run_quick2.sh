T=${1:-r02o}
timeout 300 python -m pytest tests/test_gpu_layers.py -m gpu -q 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -2
DAI_GRAPHS=0 DAI_TC_COUNTERS=1 timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 2 --quick > /dev/null 2> gpurun_out/${T}_counters.log
grep "tc counters" gpurun_out/${T}_counters.log | sort | uniq -c | sort -rn | head -${2:-8} | cut -c1-900
for rep in 1 2; do
  timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/${T}_bench_$rep.json 2>/dev/null
done
python - $T <<'PY'
import json, glob, sys
for f in sorted(glob.glob('gpurun_out/%s_bench_?.json' % sys.argv[1])):
    try:
        d = json.load(open(f))
        print(f, round(d['value'], 1), {k: round(v, 2) for k, v in d['roofline'].get('step_share_ms', {}).items()})
    except Exception as e:
        print(f, 'ERR', e)
PY

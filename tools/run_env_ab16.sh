# interleaved A/B of one environment switch at R=16 only: bash tools/run_env_ab16.sh VAR=VALUE
KV=$1
for rep in 1 2 3; do
  timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 6 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('A        ', round(d['value'],1), {k: round(v,2) for k,v in d['roofline']['step_share_ms'].items()})"
  env $KV timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 6 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('B $KV', round(d['value'],1), {k: round(v,2) for k,v in d['roofline']['step_share_ms'].items()})"
done

#!/bin/bash
# A/B of an environment switch, interleaved: tools/ab_env.sh VAR=VALUE [bench args]
SW=$1; shift
for i in 1 2 3; do
  env $SW python bench.py --steps 6 --no-cpu-baseline "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$SW  %.1f' % d['value'], d['clocks']['sm_mhz'], {k: round(v,2) for k,v in d['roofline']['step_share_ms'].items()})"
  python bench.py --steps 6 --no-cpu-baseline "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('default  %.1f' % d['value'], d['clocks']['sm_mhz'], {k: round(v,2) for k,v in d['roofline']['step_share_ms'].items()})"
done

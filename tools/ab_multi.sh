#!/bin/bash
# several environment switches against the default, interleaved twice: tools/ab_multi.sh "A=1" "B=2" ...
for i in 1 2; do
  for SW in "X=0" "$@"; do
    env $SW python bench.py --steps 6 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-22s %.1f' % ('$SW', d['value']), d['clocks']['sm_mhz'], {k: round(v,2) for k,v in d['roofline']['step_share_ms'].items()})"
  done
done

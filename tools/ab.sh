#!/bin/bash
# A/B of two builds of the library, interleaved to average out box drift: tools/ab.sh <old.so> [bench args]
OLD=$1; shift
mkdir -p gpurun_out
for i in 1 2 3; do
  DAI_B200_LIB=$PWD/$OLD python bench.py --steps 6 --no-cpu-baseline "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('old %.1f' % d['value'], d['clocks']['sm_mhz'], {k: round(v,2) for k,v in d['roofline']['step_share_ms'].items()})"
  python bench.py --steps 6 --no-cpu-baseline "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('new %.1f' % d['value'], d['clocks']['sm_mhz'], {k: round(v,2) for k,v in d['roofline']['step_share_ms'].items()})"
done

#!/bin/bash
# One GPU call: parity tests, per-layer timings / wait counters for the MMA issue orders, short bench per setting.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest.log 2>&1
tail -3 gpurun_out/pytest.log
DAI_TC_TWO_PASS=7 python -m pytest tests/test_gpu_layers.py -m gpu -x -q > gpurun_out/pytest_tp7.log 2>&1
tail -2 gpurun_out/pytest_tp7.log
for tp in none 1 7; do
  if [ $tp = none ]; then unset DAI_TC_TWO_PASS; else export DAI_TC_TWO_PASS=$tp; fi
  python tools/layer_time.py 960 > gpurun_out/lt_tp_$tp.log 2>&1
  DAI_TC_COUNTERS=1 python tools/layer_time.py 960 > gpurun_out/ltc_tp_$tp.log 2>&1
  python bench.py --steps 5 --no-cpu-baseline > gpurun_out/bench_tp_$tp.json 2> gpurun_out/bench_tp_$tp.err
done
unset DAI_TC_TWO_PASS
DAI_TC_COUNTERS=1 DAI_TC_DBG=1 python tools/layer_time.py 960 > gpurun_out/ltc_idle_epilogue.log 2>&1
grep -h "us/launch" gpurun_out/lt_tp_*.log
cat gpurun_out/bench_tp_*.json | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); print(d['value'], d['roofline']['step_share_ms'])
    except Exception as e: print('bad line', e)
"

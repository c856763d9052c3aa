#!/bin/bash
mkdir -p gpurun_out
for ch in 1024 1600 2400 3200 4800; do
  DAI_DEC_CHUNK=$ch python bench.py --steps 5 --no-cpu-baseline > gpurun_out/bench_chunk_$ch.json 2> gpurun_out/bench_chunk_$ch.err
  python - $ch <<'PY'
import json, sys
ch = sys.argv[1]
try:
    d = json.load(open("gpurun_out/bench_chunk_%s.json" % ch))
    print("chunk", ch, "rollouts/s %.1f" % d["value"], {k: round(v, 2) for k, v in d["roofline"]["step_share_ms"].items()}, d["clocks"]["sm_mhz"])
except Exception as e:
    print("chunk", ch, "failed", e)
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 1 --quick --roots 16 > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300

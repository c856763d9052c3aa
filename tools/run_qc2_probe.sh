# conv1 -> conv2: the fused kernel (in-kernel counters and ncu durations, warm caches) after the parity tests
T=${1:-r03a}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_layers.py -m gpu -q -x 2>&1 | tail -3
for f in ${2:-1}; do
  DAI_TC_FUSE_C1=$f DAI_GRAPHS=0 DAI_TC_COUNTERS=1 timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 2 --quick > /dev/null 2> gpurun_out/${T}_counters_f$f.log
  echo "== fuse_c1=$f"; grep "tc counters\] mode 2 nph 32" gpurun_out/${T}_counters_f$f.log | sort | uniq -c | sort -rn | head -3 | cut -c1-400
  DAI_TC_FUSE_C1=$f DAI_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k "regex:k_tc_conv|k_qs_conv1" -s 40 -c 60 --csv --log-file gpurun_out/${T}_list_f$f.csv python bench.py --steps 1 --quick --no-extras --no-cpu-baseline > /dev/null 2>&1
  python tools/launch_summary.py gpurun_out/${T}_list_f$f.csv
done

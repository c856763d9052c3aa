DAI_GRAPHS=0 DAI_TC_COUNTERS=1 timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 2 --quick > /dev/null 2> gpurun_out/r02n_counters.log
grep "tc counters" gpurun_out/r02n_counters.log | grep "ct2+ct3" | head -4 | cut -c1-900

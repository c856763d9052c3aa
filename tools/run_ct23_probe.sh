set -x
DAI_TC_COUNTERS=1 timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 2 --quick > /dev/null 2> gpurun_out/r02e_counters.log
grep "pair kernel" gpurun_out/r02e_counters.log | head -4
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_ct23 -s 6 -c 1 -o gpurun_out/r02e_ct23 -f python bench.py --no-extras --no-cpu-baseline --steps 1 --quick > gpurun_out/r02e_ncu.log 2>&1
tail -3 gpurun_out/r02e_ncu.log
ls -la gpurun_out/r02e_ct23.ncu-rep

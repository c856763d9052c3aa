# launch list of one sequential MCTS decision (configs[3]: 30 expansions x N=50 x depth 10), warm caches
DAI_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 2000 -c 1500 --csv --log-file gpurun_out/r03k_launches_mcts.csv python bench.py --workload mcts --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r03k_ncu_list.log 2>&1
python tools/launch_summary.py gpurun_out/r03k_launches_mcts.csv
tail -3 gpurun_out/r03k_ncu_list.log | cut -c1-300

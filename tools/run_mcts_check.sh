# simulate / planner parity tests, then configs[3] timing (sequential and device tree) and the sim kernel's duration
timeout 900 python -m pytest tests/test_planner.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
for i in 1 2; do
timeout 300 python bench.py --workload mcts --steps 6 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('sequential', round(d['value'],2), 'decisions/s', round(d['ms_per_step'],2), 'ms')"
timeout 300 python bench.py --workload mcts --leaves 16 --device-tree --steps 6 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('device tree K=16', round(d['value'],2), 'decisions/s', round(d['ms_per_step'],2), 'ms')"
done
DAI_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k k_sim_rollout -s 5 -c 20 --csv --log-file gpurun_out/r03l_sim.csv python bench.py --workload mcts --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r03l_sim.csv

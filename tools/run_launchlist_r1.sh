# launch list at R=1 (one root: 200 decoder rows per horizon step) -- where the single-root latency goes
DAI_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 600 --csv --log-file gpurun_out/r02t_launches_r1.csv python bench.py --roots 1 --steps 4 --quick --no-extras --no-cpu-baseline > gpurun_out/r02t_ncu_list.log 2>&1
python tools/launch_summary.py gpurun_out/r02t_launches_r1.csv
for g in 1 0; do DAI_GRAPHS=$g timeout 300 python bench.py --roots 1 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('graphs=$g', d['value'], d['ms_per_step'], d['e2e'])"; done

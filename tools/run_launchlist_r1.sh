# launch list at R=1 (one root: 600 decoder rows per horizon step) -- where the single-root latency goes; caches left warm
# (--cache-control none) so that the durations are those of a replayed step, not of cold weights
DAI_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 300 -c 600 --csv --log-file gpurun_out/${1:-r02w}_launches_r1.csv python bench.py --roots 1 --steps 4 --quick --no-extras --no-cpu-baseline > gpurun_out/${1:-r02w}_ncu_list.log 2>&1
python tools/launch_summary.py gpurun_out/${1:-r02w}_launches_r1.csv

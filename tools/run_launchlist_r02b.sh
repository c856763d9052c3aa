# launch list of the final state at R=16 (one time-batched rollout group sequence), warm caches
DAI_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 500 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 1 --quick --no-extras --no-cpu-baseline > gpurun_out/r02b_ncu_list.log 2>&1
python tools/launch_summary.py gpurun_out/r02b_launches.csv

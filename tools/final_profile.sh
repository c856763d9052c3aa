#!/bin/bash
# Round-2 closing profiles on one B200 (the numbers under profiles/r02_*): bench line, launch list, ncu --set full captures.
# Never a bench value: ncu serialises and replays kernels (compare SHARES, not absolutes).
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
cut -c1-200 gpurun_out/r02_bench.json
DAI_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 700 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --quick --no-extras --no-cpu-baseline > gpurun_out/r02_ncu_list.log 2>&1
DAI_GRAPHS=0 timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_tc_ct23|k_tc_fc4_pair|k_tc_conv|k_ct4_gather|k_qs_conv1|k_fc4_mask" -s 30 -c 9 -o gpurun_out/r02_decoder -f python bench.py --no-extras --no-cpu-baseline --steps 1 --quick > gpurun_out/r02_ncu_full.log 2>&1
tail -2 gpurun_out/r02_ncu_full.log | cut -c1-200
DAI_GRAPHS=0 DAI_TC_FUSE23=0 DAI_TC_FC4_PAIR=0 timeout 900 ncu --set full --clock-control none -k "regex:k_tc_conv<.*TrCt[23]|k_tc_dense<256" -s 30 -c 3 -o gpurun_out/r02_separate -f python bench.py --no-extras --no-cpu-baseline --steps 1 --quick > gpurun_out/r02_ncu_sep.log 2>&1
ls -la gpurun_out/r02_*.ncu-rep

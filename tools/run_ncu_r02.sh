# ncu --set full captures for profiles/: the fused pair kernel, and the separate decoder kernels (one launch each at 4800 rows)
set -x
DAI_GRAPHS=0 DAI_TC_FUSE23=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_ct23 -s 6 -c 1 -o gpurun_out/r02h_ct23 -f python bench.py --no-extras --no-cpu-baseline --steps 1 --quick > gpurun_out/r02h_ncu1.log 2>&1
DAI_GRAPHS=0 DAI_TC_FUSE23=0 timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_tc_conv|k_tc_dense" -s 40 -c 14 -o gpurun_out/r02h_sep -f python bench.py --no-extras --no-cpu-baseline --steps 1 --quick > gpurun_out/r02h_ncu2.log 2>&1
tail -2 gpurun_out/r02h_ncu1.log | cut -c1-200; tail -2 gpurun_out/r02h_ncu2.log | cut -c1-200
ls -la gpurun_out/r02h_*.ncu-rep

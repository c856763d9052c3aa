# ncu --set full of the final kernels (one launch each): the conv1+conv2 kernel, the pair kernels, ct1, the pixel kernel
DAI_GRAPHS=0 timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_tc_ct23|k_tc_fc4_pair|k_tc_conv|k_ct4_gather" -s 12 -c 6 -o gpurun_out/r02c_final -f python bench.py --no-extras --no-cpu-baseline --steps 1 --quick > gpurun_out/r02c_ncu.log 2>&1
tail -2 gpurun_out/r02c_ncu.log | cut -c1-200
ls -la gpurun_out/r02c_final.ncu-rep

# ncu --set full of the pair kernels after the issuer changes (one launch each, 9600 rows)
DAI_GRAPHS=0 timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_tc_ct23|k_tc_fc4_pair|k_tc_conv<.*TrCt1" -s 10 -c 3 -o gpurun_out/r02r_dec -f python bench.py --no-extras --no-cpu-baseline --steps 1 --quick > gpurun_out/r02r_ncu1.log 2>&1
tail -2 gpurun_out/r02r_ncu1.log | cut -c1-200
ls -la gpurun_out/r02r_*.ncu-rep

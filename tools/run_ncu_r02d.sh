# ncu --set full of FC4 (pair kernel), ct1 and the pixel kernel of the final build (one launch each, 16000-row chunks)
DAI_GRAPHS=0 timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_tc_fc4_pair|TrCt1|k_ct4_gather" -s 3 -c 3 -o gpurun_out/r02d_final -f python bench.py --no-extras --no-cpu-baseline --steps 1 --quick > gpurun_out/r02d_ncu.log 2>&1
tail -2 gpurun_out/r02d_ncu.log | cut -c1-200

set -x
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest "tests/test_gpu_layers.py::test_fused_ct2_ct3_pair_kernel_matches_fp32[1]" -x -q > gpurun_out/r02f_sanitize.log 2>&1
grep -v "^$" gpurun_out/r02f_sanitize.log | grep -i "=========\|error\|invalid\|trap\|passed\|failed" | head -40

# Round-2 closing measurements on one B200: full GPU test suite, sanitizer, bench (both arms), launch list, ncu full captures
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/r02_pytest.log 2>&1; tail -15 gpurun_out/r02_pytest.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 3 python tools/sanitize_case.py > gpurun_out/r02_memcheck.log 2>&1; tail -3 gpurun_out/r02_memcheck.log
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; cut -c1-300 gpurun_out/r02_bench.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; cut -c1-300 gpurun_out/r02_bench_reference.json

# A/B of the fused ct2->ct3 pair kernel against the two separate kernels (DAI_TC_FUSE23=0), after its layer tests
set -x
timeout 300 python -m pytest tests/test_gpu_layers.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02d_layers.log
cat gpurun_out/r02d_layers.log
if grep -q "passed" gpurun_out/r02d_layers.log && ! grep -q "failed" gpurun_out/r02d_layers.log; then
  timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r02d_bench_fused.json 2> gpurun_out/r02d_bench_fused.err
  DAI_TC_FUSE23=0 timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r02d_bench_unfused.json 2> gpurun_out/r02d_bench_unfused.err
  DAI_TC_COUNTERS=1 timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 2 --quick > /dev/null 2> gpurun_out/r02d_counters.log
  python -c "
import json
for f in ('fused','unfused'):
    d=json.load(open('gpurun_out/r02d_bench_%s.json'%f)); print(f, d['value'], d['roofline'].get('step_share_ms'))
"
  grep "pair kernel" gpurun_out/r02d_counters.log | head -3
  timeout 600 python -m pytest tests -m "gpu and not fullsize" -q 2>&1 | tail -8 > gpurun_out/r02d_pytest.log; cat gpurun_out/r02d_pytest.log
fi

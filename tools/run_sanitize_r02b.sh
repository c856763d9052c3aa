# memcheck of the end-to-end case (time-batched rollout, conv1-in-conv2 generators, side stream, PDL), then the chunk-size A/B
timeout 1200 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanitize_case.py > gpurun_out/r02b_memcheck.log 2>&1; tail -4 gpurun_out/r02b_memcheck.log
for rep in 1 2; do
  for c in 19200 24000; do
    DAI_DEC_CHUNK=$c timeout 300 python bench.py --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('chunk $c', round(d['value'],1), round(d['ms_per_step'],3), {k: round(v,2) for k,v in d['roofline']['step_share_ms'].items()})"
  done
done

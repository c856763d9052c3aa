#!/bin/bash
# Round-1 closing measurements: roots sweep, launch list, ncu full of the conv kernels.
mkdir -p gpurun_out
for r in 8 32 64; do
  python bench.py --steps 4 --no-cpu-baseline --roots $r 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('roots $r: %.1f rollouts/s, e2e %.1f' % (d['value'], d['e2e']['value']), d['clocks']['sm_mhz'], {k: round(v,2) for k,v in d['roofline']['step_share_ms'].items()})"
done
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
cat gpurun_out/bench_default.json | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 700 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --quick --roots 16 > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tc_conv -s 20 -c 8 -o gpurun_out/prof_conv python bench.py --steps 1 --quick --roots 16 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log | cut -c1-200
ls -la gpurun_out/prof_conv.ncu-rep

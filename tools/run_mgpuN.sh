# bench line at N GPUs (with the configs[4] sample-sharded leg): bash tools/run_mgpuN.sh N
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N > gpurun_out/r03j_bench$N.json 2> gpurun_out/r03j_bench$N.err
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/r03j_bench$N.json') if l.startswith('{')][-1]
print(round(d['value'],1), d['n_gpus'], round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1)); e=d.get('extra',{}); print({k:(round(v['value'],1), round(v['ms_per_step'],2)) for k,v in e.items() if 'value' in v})"

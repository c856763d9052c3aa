# 8 GPUs: sharded == unsharded on every rank, then the bench line with the configs[4] sample-sharded leg
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py > gpurun_out/r03f_mgpu8.log 2>&1; tail -4 gpurun_out/r03f_mgpu8.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 > gpurun_out/r03f_bench8.json 2> gpurun_out/r03f_bench8.err
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/r03f_bench8.json') if l.startswith('{')][-1]
print(round(d['value'],1), d['n_gpus'], round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1)); print(json.dumps(d.get('extra'))[:1200])"

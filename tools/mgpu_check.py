"""torchrun --nproc-per-node N tools/mgpu_check.py — on a multi-GPU box: the sample-sharded model (NCCL
all-reduce of the term sums) returns, on every rank, what the unsharded model returns."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dai_b200  # noqa: E402,F401
from dai_b200 import synthetic  # noqa: E402
from dai_b200.torchmodel import ActiveInferenceModel  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
m = ActiveInferenceModel(10, 4, 1.0, 1.0, 1.0, device="cuda:%d" % local).load_numpy_weights(synthetic.make_weights(0))
o = torch.from_numpy(synthetic.make_frames(2, 3)).repeat_interleave(4, dim=0)
m.set_rng(99, 0)
G0, terms0, po0 = m.calculate_G_repeated(o, torch.eye(4).repeat(2, 1), steps=3, samples=7)
s0 = torch.from_numpy(synthetic.make_frames(1, 5)).reshape(-1)[:40].reshape(4, 10)
m.set_rng(99, 5)
g0 = m.calculate_G(s0, torch.eye(4), samples=9)
ok = True
for native in (False, True):          # torch.distributed all-reduce of the returned sums / the library's own NCCL communicator
    m.enable_sample_sharding(native=native)
    m.set_rng(99, 0)
    G1, terms1, po1 = m.calculate_G_repeated(o, torch.eye(4).repeat(2, 1), steps=3, samples=7)
    ok = ok and torch.allclose(G0, G1, rtol=1e-5, atol=1e-4) and torch.allclose(po0, po1, atol=1e-6)
    for a, b in zip(terms0, terms1):
        ok = ok and torch.allclose(a, b, rtol=1e-5, atol=2e-3)
    m.set_rng(99, 5)
    g1 = m.calculate_G(s0, torch.eye(4), samples=9)
    ok = ok and torch.allclose(g0[0], g1[0], rtol=1e-5, atol=1e-4) and torch.equal(g0[2], g1[2]) and torch.equal(g0[4], g1[4])
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if dist.get_rank() == 0:
    print("mgpu_check world=%d: %s  G=%s" % (dist.get_world_size(), "OK" if flag.item() else "MISMATCH", G1.tolist()))
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)

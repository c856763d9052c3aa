"""World-size-2 gloo test of the MC-sample sharding protocol on CPU (SURVEY.md §8 e): contiguous sample
shards, the last-sample recompute on every rank, ONE all-reduce of the (4,B) float64 partial sums, combine.
Each rank evaluates its shard with the oracle (test infrastructure); the product pieces under test are
dai_b200.sharding.{shard_range, allreduce_term_sums, combine_sums}."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, samples, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    import dai_b200  # noqa: F401
    from dai_b200.sharding import shard_range, allreduce_term_sums, combine_sums
    from dai_b200 import synthetic
    from oracle import efe_oracle as O
    W = O.to_torch(synthetic.make_weights(0))
    s0 = torch.from_numpy(np.random.default_rng(5).standard_normal((4, 10)).astype(np.float32))
    pi = torch.eye(4)
    j0, j1 = shard_range(samples, rank, world)
    sums, last = O.calculate_G_shard(W, s0, pi, samples, j0, j1, O.PhiloxNoise(4242))
    allreduce_term_sums(sums)
    G, t0, t1, t2 = combine_sums(sums, samples)
    # every rank must hold the same result and the same carry
    gathered = [torch.zeros_like(G) for _ in range(world)]
    dist.all_gather(gathered, G)
    carry = [torch.zeros_like(last[0]) for _ in range(world)]
    dist.all_gather(carry, last[0].contiguous())
    if rank == 0:
        torch.save({"G": G, "t0": t0, "t1": t1, "t2": t2, "gathered": gathered, "carry": carry, "ps1": last[0]}, out_path)
    dist.destroy_process_group()


@pytest.mark.parametrize("samples", [3, 4])
def test_two_rank_sample_sharding_matches_unsharded(tmp_path, samples):
    sys.path.insert(0, ROOT)
    import dai_b200  # noqa: F401
    from dai_b200 import synthetic
    from oracle import efe_oracle as O
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), samples, out), nprocs=2, join=True)
    got = torch.load(out)
    W = O.to_torch(synthetic.make_weights(0))
    s0 = torch.from_numpy(np.random.default_rng(5).standard_normal((4, 10)).astype(np.float32))
    G, terms, ps1, ps1_mean, po1 = O.calculate_G(W, s0, torch.eye(4), samples, O.PhiloxNoise(4242))
    assert torch.allclose(got["G"], G, rtol=1e-5, atol=1e-4)
    assert torch.allclose(got["t0"], terms[0], rtol=1e-5, atol=1e-5)
    assert torch.allclose(got["t1"], terms[1], rtol=1e-5, atol=1e-5)
    assert torch.allclose(got["t2"], terms[2], rtol=0, atol=2e-3)
    assert torch.equal(got["gathered"][0], got["gathered"][1])
    assert torch.equal(got["carry"][0], got["carry"][1]) and torch.equal(got["ps1"], ps1)

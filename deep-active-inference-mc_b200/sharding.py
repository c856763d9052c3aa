"""MC-sample sharding of the EFE rollout over GPUs (SURVEY.md §8 e).

The N samples of every calculate_G step are independent (src/torchmodel.py:273,287), so rank r
of W evaluates the contiguous slice shard_range(N, r, W) for all B rows with replicated
weights.  Noise is addressed by the GLOBAL sample index, so the union over ranks is exactly
the single-GPU evaluation.  The two cross-sample dependencies — the step carry s0 <- ps1 and
loop 2b's reparameterize(ps1_mean, ps1_logvar), both of the globally last sample
(:243,266,291,300) — are resolved without communication: every rank recomputes that one
transition under the same key.  What remains is one all-reduce(sum) of the (4,B) float64
partial sums of term0, term1, term2_1, term2_2 per rollout.
"""


def shard_range(samples, rank, world):
    """Contiguous [begin, end) of `samples` for `rank` of `world`; the first samples % world ranks hold one more."""
    if world <= 0 or not (0 <= rank < world) or samples < 0:
        raise ValueError("bad shard request samples=%d rank=%d world=%d" % (samples, rank, world))
    base, rem = divmod(samples, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def allreduce_term_sums(sums, group=None):
    """The one collective of a sharded rollout: in-place sum of the (4,B) float64 partial term sums."""
    import torch.distributed as dist
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


def combine_sums(sums, samples):
    """(4,B) float64 sums of term0, term1, term2_1, term2_2 -> (G, term0, term1, term2) float32, as
    src/torchmodel.py:282-298 finishes them.  Host-side twin of dai_combine (which the model uses for device
    tensors); used where the sums live on the CPU (gloo tests)."""
    import torch
    t = (sums / float(samples)).to(torch.float32)
    t2 = t[2] - t[3]
    return -t[0] + t[1] + t2, t[0], t[1], t2

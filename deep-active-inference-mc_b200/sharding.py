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

"""Statistical validation of the single-product fast mode (DAI_PREC_BF16X1; SURVEY.md §4 "statistical (GPU)").

bf16x1 issues one bf16 product per MAC instead of three, so it misses the 1e-4 per-evaluation parity bar (it is NOT
the parity mode and no parity test runs in it).  What it must preserve is the Monte-Carlo estimate itself: over
>= 10^4 samples, the mean and the sample-to-sample spread of every EFE term have to sit inside the confidence interval
of the parity mode's estimate — which is pinned to the reference/oracle at 1e-4 by tests/test_gpu_parity.py — on the
same keyed noise AND on independent noise (a different seed), for the default and the saturating weight set."""
import numpy as np
import pytest
import torch

import cases

pytestmark = pytest.mark.gpu

N, SHARD = 12000, 500          # 24 shards of 500 samples: shard means give the standard error of the MC estimate


def _shard_means(kind, precision, seed):
    from dai_b200.torchmodel import ActiveInferenceModel
    m = ActiveInferenceModel(10, 4, 1.0, 1.0, 1.0, precision=precision, device="cuda:0").load_numpy_weights(cases.weights_for(kind))
    m._sync()
    eng = m._engine
    s0 = torch.from_numpy(np.random.default_rng(6).standard_normal((4, 10)).astype(np.float32)).cuda()
    pi = torch.eye(4, device="cuda")
    out = []
    for j0 in range(0, N, SHARD):
        eng.set_rng(seed, 0)
        part = eng.calculate_G(s0, pi, N, shard=(j0, j0 + SHARD), want_po1=False)
        out.append((part["sums"] / SHARD).cpu().numpy())         # (4, B): term0, term1, term2_1, term2_2 means of this shard
    return np.stack(out)                                          # (shards, 4, B)


@pytest.mark.parametrize("kind", ["w0", "w0s"])
def test_bf16x1_estimates_sit_inside_the_parity_modes_confidence_interval(kind):
    ref = _shard_means(kind, "bf16x3", 4242)
    same = _shard_means(kind, "bf16x1", 4242)                     # same noise: isolates the arithmetic
    other = _shard_means(kind, "bf16x1", 977)                     # independent noise: the estimator as a user sees it
    k = ref.shape[0]
    mean_ref, se_ref = ref.mean(0), ref.std(0, ddof=1) / np.sqrt(k)
    G_ref = -mean_ref[0] + mean_ref[1] + (mean_ref[2] - mean_ref[3])
    # (1) same noise: the bias of the fast arithmetic is far inside the MC standard error, and tiny against |G|
    bias = same.mean(0) - mean_ref
    assert np.all(np.abs(bias) <= 0.5 * se_ref + 1e-6), (bias, se_ref)
    G_same = -same.mean(0)[0] + same.mean(0)[1] + (same.mean(0)[2] - same.mean(0)[3])
    assert np.all(np.abs(G_same - G_ref) <= 2e-3 * np.abs(G_ref)), (G_same, G_ref)
    # (2) independent noise: means agree within the combined 4-sigma interval, spreads within a factor ~1.6 (F-test, 23 dof)
    se_other = other.std(0, ddof=1) / np.sqrt(k)
    z = np.abs(other.mean(0) - mean_ref) / np.sqrt(se_ref ** 2 + se_other ** 2 + 1e-12)
    assert np.all(z < 4.5), z
    ratio = (other.var(0, ddof=1) + 1e-12) / (ref.var(0, ddof=1) + 1e-12)
    assert np.all((ratio > 0.3) & (ratio < 3.3)), ratio
    # (3) same noise, shard by shard: the two modes track each other sample block by sample block
    assert np.all(np.abs(same - ref) <= 5e-3 * np.abs(ref) + 1e-4)

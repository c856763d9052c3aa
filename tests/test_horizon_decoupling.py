"""What the time-batched rollout schedule rests on, checked on the oracle — the restatement of
src/torchmodel.py:227-300 that is pinned bit for bit against the real reference: the horizon steps of
calculate_G_repeated are chained ONLY through the transition net (the state a step starts from is the previous step's
ps1 — or its mean — of the LAST Monte-Carlo sample), so (1) the latent chain of all steps can be computed first, from Ps
alone, and (2) every step's EFE terms can then be evaluated from its start state in any order.  CPU, keyed noise."""
import numpy as np
import pytest
import torch

import cases
from oracle import efe_oracle as O


@pytest.mark.parametrize("calc_mean,four", [(False, True), (True, False), (True, True)])
def test_horizon_steps_are_chained_only_through_the_transition_net(calc_mean, four):
    W = O.to_torch(cases.weights_for("w0"))
    rng = np.random.default_rng(2)
    B, T, N = 4, 3, 2
    o = torch.from_numpy(rng.random((B, 1, 64, 64), dtype=np.float32))
    pi = torch.eye(4)
    key = cases.SEED + 17

    trace = []
    sum_G, sum_terms, po1 = O.calculate_G_repeated(W, o, pi, T, calc_mean, N, O.PhiloxNoise(key), four=four, trace=trace)

    # (1) the latent chain from the transition net alone: the loop-2a transition of the last sample of every step
    nz = O.PhiloxNoise(key)
    nz.at(0, 0)
    m0, lv0 = O.qs_forward(W, o, nz, O.SITES["QS_ROOT"])
    s = m0 if calc_mean else O.reparameterize(m0, lv0, nz, O.SITES["QS_ROOT"] + 3)
    mean_variant = four and calc_mean
    starts = []
    for t in range(T):
        starts.append(s)
        nz.at(t, 0 if mean_variant else N - 1)
        ps1, ps1_mean, _ = O.ps_forward_with_sample(W, pi, s, nz, O.SITES["PS_A"])
        s = ps1_mean if calc_mean else ps1
    for t in range(T):
        assert torch.equal(starts[t], trace[t]["s0"]), t

    # (2) the steps evaluated from those states in REVERSE order give the same per-step G and terms
    for t in reversed(range(T)):
        nz2 = O.PhiloxNoise(key)
        if mean_variant:
            G, terms, _, _ = O.calculate_G_mean(W, starts[t], pi, nz2, step=t)
        else:
            G, terms, _, _, _ = O.calculate_G(W, starts[t], pi, N, nz2, step=t)
        assert torch.equal(G, trace[t]["G"]), t
        for a, b in zip(terms, trace[t]["terms"]):
            assert torch.equal(a, b), t
    assert torch.equal(sum_G, sum(tr["G"] for tr in trace))

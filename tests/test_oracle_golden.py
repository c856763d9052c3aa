"""The oracle against the fixtures written from the real reference
(tests/golden/make_golden.py): bit-exact on the authoring image, and inside the parity
tolerance anywhere (a different BLAS thread count may reorder fp32 sums)."""
import numpy as np
import pytest

import cases
from oracle import efe_oracle as O

_models = {}


def _oracle(kind):
    if kind not in _models:
        _models[kind] = O.OracleModel(cases.weights_for(kind), seed=cases.SEED)
    return _models[kind]


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_oracle_matches_reference_fixture(name, golden):
    got = cases.run_case(name, _oracle(cases.CASES[name][0]))
    assert set(got) == set(golden[name])
    assert cases.compare(name, got, golden[name], rtol=2e-6, atol=2e-6) == []


def test_reward_closed_form():
    """check_reward is linear in o with the four fp32 constants of SURVEY.md §8 a6 (D12)."""
    import torch
    o = torch.rand(3, 1, 64, 64, generator=torch.Generator().manual_seed(0))
    d = np.float32(1e-5)
    one = np.float32(1.0)
    a_top, b_top = np.log(np.float32(d + one)), np.log(np.float32(np.float32(d + one) - one))
    a_bot, b_bot = np.log(d), np.log(np.float32(d + one))
    x = o.numpy()[:, 0].astype(np.float64)
    top = x[:, :32] * a_top + (1 - x[:, :32]) * b_top
    bot = x[:, 32:] * a_bot + (1 - x[:, 32:]) * b_bot
    want = 10.0 * (top.sum((1, 2)) + bot.sum((1, 2))) / 4096.0
    assert np.allclose(O.check_reward(o).numpy(), want, rtol=2e-6)


def test_eval_mode_is_identity_dropout():
    m = O.OracleModel(cases.weights_for("w0"), seed=1, training=False)
    import torch
    s = torch.zeros(2, 10)
    m.set_rng(1, 0)
    a = m.model_down.decoder(s)
    m.set_rng(2, 5)
    b = m.model_down.decoder(s)
    assert torch.equal(a, b)

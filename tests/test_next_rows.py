"""SURVEY.md §8 (f) "next" rows built so far: f1 batched many-roots action selection, f3 checkpoint round trip."""
import os
import sys

import numpy as np
import pytest
import torch

import cases
from oracle import efe_oracle as O

REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference not mounted (authoring container only)")
def test_oracle_softmax_matches_reference_util():
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from src.util import softmax_multi_with_log          # the reference's own function
    G = torch.from_numpy(np.random.default_rng(0).normal(30, 3, size=24).astype(np.float32))
    SM_ref, log_ref = softmax_multi_with_log(-G.numpy(), 4)
    SM, logSM, choices = O.select_actions(G, 10.0, O.PhiloxNoise(9))
    assert np.array_equal(SM, SM_ref) and np.array_equal(logSM, log_ref)
    assert choices.shape == (6,) and choices.min() >= 0 and choices.max() <= 3


def test_oracle_choice_frequencies_follow_ppi():
    G = torch.tensor([1.0, 5.0, 9.0, 30.0] * 4000)
    SM, _, choices = O.select_actions(G, 10.0, O.PhiloxNoise(123))
    freq = np.bincount(choices, minlength=4) / len(choices)
    assert np.allclose(freq, SM[0], atol=0.02)


@pytest.mark.gpu
def test_select_actions_matches_oracle():
    from dai_b200.torchmodel import ActiveInferenceModel
    import dai_b200.synthetic as syn
    w = cases.weights_for("w0")
    gpu = ActiveInferenceModel(10, 4, 1.0, 1.0, 1.0, device="cuda:0").load_numpy_weights(w)
    ora = O.OracleModel(w, seed=31)
    frames = torch.from_numpy(syn.make_frames(5, 40))
    gpu.set_rng(31, 0)
    cg, pg, lg, Gg, _ = gpu.select_actions(frames, steps=2, samples=2)
    co, po, lo, Go, _ = ora.select_actions(frames, steps=2, samples=2)
    assert torch.allclose(Gg.cpu(), Go, rtol=1e-4)
    assert torch.allclose(pg, po, rtol=1e-3, atol=1e-5) and torch.allclose(lg, lo, rtol=1e-3, atol=1e-3)
    assert torch.equal(cg.to(torch.int32), co.to(torch.int32))
    # the kernel alone on a fixed G: exact categorical draws, many roots
    G = torch.from_numpy(np.random.default_rng(1).normal(30, 3, size=4 * 3000).astype(np.float32))
    gpu._engine.set_rng(77, 5)
    P, L, C = gpu._engine.select_actions(G.cuda(), 10.0)
    SM, logSM, ch = O.select_actions(G, 10.0, O.PhiloxNoise(77 + 5))
    assert np.allclose(P.cpu().numpy(), SM, rtol=1e-5, atol=1e-7) and np.allclose(L.cpu().numpy(), logSM, rtol=1e-5, atol=1e-5)
    assert (C.cpu().numpy() != ch).mean() < 1e-3          # a draw can only flip where u*total sits within an ulp of a CDF edge


@pytest.mark.gpu
def test_checkpoint_round_trip(tmp_path):
    """save_weights / load_weights write and read the reference's three files (src/torchmodel.py:167-177)."""
    from dai_b200.torchmodel import ActiveInferenceModel
    a = ActiveInferenceModel(10, 4, 1.0, 1.0, 1.0, device="cuda:0").load_numpy_weights(cases.weights_for("w0"))
    a.save_weights(str(tmp_path))
    assert sorted(os.listdir(tmp_path)) == ["checkpoint_down.pth", "checkpoint_mid.pth", "checkpoint_top.pth"]
    sd = torch.load(str(tmp_path / "checkpoint_down.pth"))
    assert tuple(sd["qs_net.9.weight"].shape) == (256, 576) and tuple(sd["po_net.9.weight"].shape) == (16384, 256)
    b = ActiveInferenceModel(10, 4, 1.0, 1.0, 1.0, device="cuda:0")
    b.load_weights(str(tmp_path))
    s = torch.zeros(2, 10, device="cuda")
    a.set_rng(3, 0); b.set_rng(3, 0)
    assert torch.equal(a.model_down.decoder(s), b.model_down.decoder(s))
    stats = {"var_beta_s": [0.5], "var_gamma": [], "var_beta_o": [2.0]}
    a.save_all(str(tmp_path), stats)
    st, opt = b.load_all(str(tmp_path))
    assert float(b.beta_s) == 0.5 and float(b.beta_o) == 2.0 and opt == {}

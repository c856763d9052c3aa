"""Parity of the CUDA path (through the drop-in model -> ctypes -> C ABI) against the oracle
and against the fixtures written from the real reference.  Tolerances: cases.compare
(1e-4 relative on G/term0/term1, term2 against |G|, images/latents rtol 1e-4 + atol 1e-5)."""
import numpy as np
import pytest
import torch

import cases

pytestmark = pytest.mark.gpu

_models, _oracles = {}, {}


def _model(kind, precision):
    from dai_b200.torchmodel import ActiveInferenceModel
    key = (kind, precision)
    if key not in _models:
        m = ActiveInferenceModel(10, 4, 1.0, 1.0, 1.0, precision=precision, device="cuda:0")
        _models[key] = m.load_numpy_weights(cases.weights_for(kind))
    return _models[key]


def _oracle(kind):
    from oracle import efe_oracle as O
    if kind not in _oracles:
        _oracles[kind] = O.OracleModel(cases.weights_for(kind), seed=cases.SEED)
    return _oracles[kind]


@pytest.mark.parametrize("precision", ["fp32_simt", "bf16x3"])
@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_case_matches_reference_fixture_and_oracle(name, precision, golden):
    kind = cases.CASES[name][0]
    got = cases.run_case(name, _model(kind, precision), dev="cuda:0")
    assert cases.compare(name, got, golden[name]) == []
    ref = cases.run_case(name, _oracle(kind))
    assert cases.compare(name, got, ref) == []


@pytest.mark.parametrize("precision", ["fp32_simt", "bf16x3"])
def test_sample_shards_sum_to_the_full_evaluation(precision):
    """Single-process emulation of W ranks: partial sums over sample shards add up to the
    unsharded call, and the last-sample outputs agree on every shard (SURVEY.md §8 e)."""
    from dai_b200.sharding import shard_range
    m = _model("w0", precision)
    eng = m._engine
    m._sync()
    s0 = torch.from_numpy(np.random.default_rng(3).standard_normal((8, 10)).astype(np.float32)).cuda()
    pi = torch.eye(4, device="cuda").repeat(2, 1)
    N = 7
    eng.set_rng(77, 0)
    full = eng.calculate_G(s0, pi, N)
    for W in (2, 3):
        sums = torch.zeros_like(full["sums"])
        for r in range(W):
            eng.set_rng(77, 0)
            part = eng.calculate_G(s0, pi, N, shard=shard_range(N, r, W))
            sums += part["sums"]
            for k in ("ps1", "ps1_mean", "ps1_logvar", "po1"):
                assert torch.equal(part[k], full[k]), (k, W, r)
        assert torch.allclose(sums, full["sums"], rtol=1e-6, atol=1e-6)
        G, t0, t1, t2 = eng.combine(sums, N)
        assert torch.allclose(G, full["G"], rtol=1e-5, atol=1e-4)


def test_eval_mode_disables_dropout():
    m = _model("w0", "fp32_simt")
    s = torch.zeros(3, 10, device="cuda")
    try:
        m.model_down.eval(); m.model_mid.eval()
        m.set_rng(1, 0)
        a = m.model_down.decoder(s)
        m.set_rng(2, 9)
        b = m.model_down.decoder(s)
        assert torch.equal(a, b)
        from oracle import efe_oracle as O
        ora = O.OracleModel(cases.weights_for("w0"), seed=1, training=False)
        ref = ora.model_down.decoder(torch.zeros(3, 10))
        assert torch.allclose(a.cpu(), ref, rtol=1e-4, atol=1e-5)
    finally:
        m.model_down.train(); m.model_mid.train()


def test_weight_updates_are_picked_up():
    m = _model("w0", "fp32_simt")
    s = torch.zeros(2, 10, device="cuda")
    m.set_rng(5, 0)
    a = m.model_down.decoder(s)
    with torch.no_grad():
        m.model_down.po_net[19].bias.add_(0.5)
    m.set_rng(5, 0)
    b = m.model_down.decoder(s)
    with torch.no_grad():
        m.model_down.po_net[19].bias.sub_(0.5)
    m.set_rng(5, 0)
    c = m.model_down.decoder(s)
    assert not torch.equal(a, b) and torch.allclose(a, c, atol=1e-6)


def test_abi_rejects_bad_arguments_without_crashing():
    """Every entry point returns a negative code (never throws / aborts) on bad input (include/dai_b200.h)."""
    import ctypes
    from dai_b200 import engine
    m = _model("w0", "bf16x3")
    m._sync()
    eng, lib = m._engine, m._engine.lib
    st = eng._stream()
    s0 = torch.zeros(4, 10, device="cuda")
    pi = torch.eye(4, device="cuda")
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    assert lib.dai_calculate_G(eng.h, p(s0), p(pi), 4, 0, 0, 0, None, None, None, None, None, None, None, None, None, st) == -1
    assert lib.dai_calculate_G(eng.h, p(s0), p(pi), 4, 5, 3, 9, None, None, None, None, None, None, None, None, None, st) == -1
    assert lib.dai_rollout(eng.h, None, None, 4, 1, 1, 0, 0, 0, 1, None, None, None, None, None, None, st) == -1
    assert lib.dai_rollout(eng.h, p(torch.zeros(3, 4096, device="cuda")), None, 3, 1, 1, 0, 0, 0, 1, None, None, None, None, None, None, st) == -1
    assert b"B % 4" in lib.dai_last_error(eng.h)
    assert lib.dai_set_precision(eng.h, 7) == -1
    g = ctypes.c_float()
    assert lib.dai_mcts_simulate(eng.h, p(s0), 0, 0, ctypes.byref(g), p(pi), p(pi), st) == -1
    with pytest.raises(engine.DaiError):
        eng.transition(torch.eye(4, device="cuda"), torch.zeros(5, 10, device="cuda"))
    # a fresh handle without weights refuses to compute
    fresh = engine.Engine(device="cuda:0")
    with pytest.raises(engine.DaiError, match="weights"):
        fresh.decode(torch.zeros(1, 10))
    fresh.close()
    # and the handle still works afterwards
    assert torch.isfinite(m.model_down.decoder(torch.zeros(1, 10))).all()


def test_large_batch_and_many_samples_chunking():
    """More decoder rows than one activation chunk and more encoder rows than one chunk: chunk boundaries must not
    change the result — a big batched call against the same rows evaluated separately, and a handle with small
    decoder chunks (env DAI_DEC_CHUNK, read at dai_create: 7 chunks of 1056 rows, odd row counts for the CTA-pair
    kernels) against the default one-chunk handle, bit for bit."""
    import os
    from dai_b200.torchmodel import ActiveInferenceModel
    m = _model("w0", "bf16x3")
    eng = m._engine
    m._sync()
    os.environ["DAI_DEC_CHUNK"] = "1031"
    try:
        small = ActiveInferenceModel(10, 4, 1.0, 1.0, 1.0, precision="bf16x3", device="cuda:0").load_numpy_weights(cases.weights_for("w0"))
    finally:
        del os.environ["DAI_DEC_CHUNK"]
    small._sync()
    s0 = torch.from_numpy(np.random.default_rng(11).standard_normal((8, 10)).astype(np.float32)).cuda()
    pi = torch.eye(4, device="cuda").repeat(2, 1)
    eng.set_rng(5, 0)
    big = eng.calculate_G(s0, pi, 300)             # 3*300*8 = 7200 decoder rows, 2400 encoder rows
    eng.set_rng(5, 0)
    a = eng.calculate_G(s0, pi, 300, shard=(0, 100))
    eng.set_rng(5, 0)
    b = eng.calculate_G(s0, pi, 300, shard=(100, 300))
    assert torch.allclose(a["sums"] + b["sums"], big["sums"], rtol=1e-9, atol=1e-6)
    assert torch.equal(a["ps1"], big["ps1"]) and torch.equal(b["po1"], big["po1"])
    small._engine.set_rng(5, 0)
    chunked = small._engine.calculate_G(s0, pi, 300)
    assert torch.equal(chunked["sums"], big["sums"]) and torch.equal(chunked["po1"], big["po1"]) and torch.equal(chunked["G"], big["G"])


# ---- BASELINE.json configs at full size ----------------------------------------------------------------------

def _frames4(seed):
    import dai_b200.synthetic as syn
    return torch.from_numpy(syn.make_frames(1, seed)).repeat(4, 1, 1, 1)


def test_config2_full_size_matches_the_oracle():
    """configs[1]: N=50 samples, T=10 steps, one root (4 action rows), against the CPU oracle on the same keyed noise."""
    m = _model("w0", "bf16x3")
    m.set_rng(cases.SEED, 0)
    got = cases._g4(m, "cuda:0", 10, 50, False, 21)
    ora = _oracle("w0")
    ora.set_rng(cases.SEED, 0)
    with torch.no_grad():
        ref = cases._g4(ora, "cpu", 10, 50, False, 21)
    got = {k: v.detach().cpu().numpy() for k, v in got.items()}
    ref = {k: v.detach().cpu().numpy() for k, v in ref.items()}
    assert cases.compare("config2", got, ref) == []


def test_config3_many_samples_slice_matches_the_oracle_and_full_size_is_shard_additive():
    """configs[2]: N=200, T=10.  Parity with the oracle on a T=2 slice (the oracle needs ~6 s per step at N=200); at the
    full horizon the size-independent property: sample shards add up to the unsharded rollout, bit-equal carries."""
    m = _model("w0", "bf16x3")
    m.set_rng(cases.SEED, 0)
    got = cases._g4(m, "cuda:0", 2, 200, False, 22)
    ora = _oracle("w0")
    ora.set_rng(cases.SEED, 0)
    with torch.no_grad():
        ref = cases._g4(ora, "cpu", 2, 200, False, 22)
    assert cases.compare("config3", {k: v.detach().cpu().numpy() for k, v in got.items()},
                         {k: v.detach().cpu().numpy() for k, v in ref.items()}) == []
    eng = m._engine
    o = _frames4(23).cuda()
    eng.set_rng(9, 0)
    full = eng.rollout(o, None, 10, 200, four=True)
    sums = torch.zeros_like(full["sums"])
    for j0, j1 in ((0, 70), (70, 71), (71, 200)):
        eng.set_rng(9, 0)
        part = eng.rollout(o, None, 10, 200, four=True, shard=(j0, j1))
        sums += part["sums"]
        assert torch.equal(part["po1"], full["po1"])
    assert torch.allclose(sums, full["sums"], rtol=1e-9, atol=1e-6)
    eng.set_rng(9, 0)
    again = eng.rollout(o, None, 10, 200, four=True)
    assert torch.equal(again["G"], full["G"]) and torch.equal(again["sums"], full["sums"])       # deterministic


def test_config5_eight_sample_shards_add_up():
    """configs[4]: N=800 samples over 8 ranks (100 each), T=15: the eight partial sums — what the one all-reduce adds —
    equal the unsharded evaluation."""
    from dai_b200.sharding import shard_range
    m = _model("w0", "bf16x3")
    eng = m._engine
    m._sync()
    o = _frames4(24).cuda()
    eng.set_rng(3, 0)
    full = eng.rollout(o, None, 15, 800, four=True)
    sums = torch.zeros_like(full["sums"])
    for r in range(8):
        eng.set_rng(3, 0)
        part = eng.rollout(o, None, 15, 800, four=True, shard=shard_range(800, r, 8))
        sums += part["sums"]
        assert torch.equal(part["po1"], full["po1"])
    assert torch.allclose(sums, full["sums"], rtol=1e-9, atol=1e-6)
    G, t0, t1, t2 = eng.combine(sums, 800)
    assert torch.allclose(G, full["G"], rtol=1e-6, atol=1e-5)
    assert torch.isfinite(G).all() and float(G.min()) > 0


# ---- edge cases ------------------------------------------------------------------------------------------------

def test_ragged_batches_and_single_sample_match_the_oracle():
    """B not a multiple of 4 (explicit pi), one sample, one step; and eval mode on the tensor-core path."""
    from oracle import efe_oracle as O
    m = _model("w0", "bf16x3")
    rng = np.random.default_rng(5)
    import dai_b200.synthetic as syn
    o = torch.from_numpy(syn.make_frames(5, 31))
    pi = torch.eye(4)[torch.tensor([0, 3, 1, 2, 2])]
    ora = O.OracleModel(cases.weights_for("w0"), seed=8)
    m.set_rng(8, 0)
    G, terms, po1 = m.calculate_G_repeated(o, pi, steps=1, samples=1)
    with torch.no_grad():
        Go, to, poo = ora.calculate_G_repeated(o, pi, steps=1, samples=1)
    got = dict(G=G, t0=terms[0], t1=terms[1], t2=terms[2], po1=po1)
    ref = dict(G=Go, t0=to[0], t1=to[1], t2=to[2], po1=poo)
    assert cases.compare("ragged", {k: v.detach().cpu().numpy() for k, v in got.items()},
                         {k: v.detach().cpu().numpy() for k, v in ref.items()}) == []
    # eval mode: dropout is identity in all three nets, the normals are still drawn
    s0 = torch.from_numpy(rng.standard_normal((7, 10)).astype(np.float32))
    pi7 = torch.eye(4)[torch.tensor([0, 1, 2, 3, 0, 1, 2])]
    try:
        for net in (m.model_down, m.model_mid, m.model_top):
            net.eval()
        ora_eval = O.OracleModel(cases.weights_for("w0"), seed=8, training=False)
        m.set_rng(8, 3); ora_eval.set_rng(8, 3)
        Gg, tg, ps1, mu, po = m.calculate_G(s0, pi7, samples=3)
        with torch.no_grad():
            Go, to, ps1o, muo, poo = ora_eval.calculate_G(s0, pi7, samples=3)
        got = dict(G=Gg, t0=tg[0], t1=tg[1], t2=tg[2], ps1=ps1, mean=mu, po1=po)
        ref = dict(G=Go, t0=to[0], t1=to[1], t2=to[2], ps1=ps1o, mean=muo, po1=poo)
        assert cases.compare("eval", {k: v.detach().cpu().numpy() for k, v in got.items()},
                             {k: v.detach().cpu().numpy() for k, v in ref.items()}) == []
    finally:
        for net in (m.model_down, m.model_mid, m.model_top):
            net.train()


def test_new_entry_points_reject_bad_arguments():
    import ctypes
    from dai_b200 import engine
    m = _model("w0", "bf16x3")
    m._sync()
    eng, lib = m._engine, m._engine.lib
    st = eng._stream()
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    s = torch.zeros(4, 10, device="cuda")
    g = (ctypes.c_float * 4)()
    pi0, qpi = torch.zeros(4 * 3, 4, device="cuda"), torch.zeros(4, 4, device="cuda")
    assert lib.dai_mcts_simulate_batch(eng.h, p(s), 0, 3, 0, g, p(pi0), p(qpi), st) == -1
    assert lib.dai_mcts_simulate_batch(eng.h, p(s), 4, 0, 0, g, p(pi0), p(qpi), st) == -1
    assert lib.dai_mcts_simulate_batch(eng.h, None, 4, 3, 0, g, p(pi0), p(qpi), st) == -1
    prm = engine.DaiMctsParams(1.0, 2.0, 4, 1, 2, 1, 0, 1, 33)          # leaves > 32
    res = engine.DaiMctsResult()
    path = (ctypes.c_int32 * 64)()
    frame = torch.zeros(4096, device="cuda")
    assert lib.dai_mcts_plan(eng.h, p(frame), None, ctypes.byref(prm), ctypes.byref(res), path, None, None, None, st) == -1
    assert b"leaves" in lib.dai_last_error(eng.h)
    prm.leaves = 4
    assert lib.dai_mcts_plan(eng.h, None, None, ctypes.byref(prm), ctypes.byref(res), path, None, None, None, st) == -1
    # repeats = 0: only the root expansion; the decision is the best first action
    prm.repeats = 0
    assert lib.dai_mcts_plan(eng.h, p(frame), None, ctypes.byref(prm), ctypes.byref(res), path, None, None, None, st) == 0
    assert res.repeats_done == 0 and res.path_len == 1 and 0 <= path[0] < 4
    # frame producer without a sprite table, then with inconsistent sizes
    fresh = engine.Engine(device="cuda:0")
    r = torch.zeros(2, device="cuda")
    s7 = torch.zeros(2, 7, device="cuda")
    o = torch.zeros(2, 4096, device="cuda")
    assert lib.dai_frames_render(fresh.h, p(s7), 7, p(r), 2, 0, p(o), None, st) == -1
    imgs = torch.zeros(6, 4096, dtype=torch.uint8)
    sizes = (ctypes.c_int32 * 6)(1, 1, 1, 1, 2, 2)                      # multiplies to 4, table has 6
    assert lib.dai_frames_set_sprites(fresh.h, ctypes.c_void_p(imgs.data_ptr()), 6, sizes, st) == -1
    fresh.close()
    assert torch.isfinite(m.model_down.decoder(torch.zeros(1, 10))).all()


def test_encoder_batches_larger_than_one_chunk_match_the_oracle():
    """More than 4096 image rows in ONE slot (dai_encode / the root encode of a rollout with > 1024 roots / the
    trajectory rows of a large simulation batch): the launch is split by rows and the noise row index must run on."""
    from oracle import efe_oracle as O
    import dai_b200.synthetic as syn
    m = _model("w0", "bf16x3")
    B = 4500
    o = torch.from_numpy(syn.make_frames(60, 41)).repeat(75, 1, 1, 1)[:B]
    ora = O.OracleModel(cases.weights_for("w0"), seed=19)
    m.set_rng(19, 0)
    s, mean, logvar = m.model_down.encoder_with_sample(o)
    with torch.no_grad():
        so, meano, logvaro = ora.model_down.encoder_with_sample(o)
    got = dict(s=s, mean=mean, logvar=logvar)
    ref = dict(s=so, mean=meano, logvar=logvaro)
    assert cases.compare("big_encode", {k: v.detach().cpu().numpy() for k, v in got.items()},
                         {k: v.detach().cpu().numpy() for k, v in ref.items()}) == []

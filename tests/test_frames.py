"""the frame producer (src/game_environment.py:39-66) — oracle vs the reference's own Game methods,
oracle vs golden fixture, CUDA kernel vs oracle (bit-exact: byte/index work)."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import frames_oracle as F

REF = "/root/reference"
GOLD = os.path.join(os.path.dirname(__file__), "golden", "frames_golden.npz")
SIZES = (1, 3, 2, 4, 8, 8)                       # a small table with dSprites' latent structure: 1536 sprites


def _games(n, seed, sizes=SIZES):
    rng = np.random.default_rng(seed)
    s = np.zeros((n, 7), dtype=np.float32)
    for i, m in enumerate(sizes):
        s[:, i] = rng.integers(0, m, size=n)
    s[:, 6] = rng.uniform(-10, 10, size=n)
    r = rng.uniform(-1, 1, size=n).astype(np.float32)
    r[0], r[1 % n] = 0.0, -1.0
    if n > 2:
        r[2] = 1.0
    return s, r


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference not mounted (authoring container only)")
def test_oracle_matches_the_reference_game_methods():
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from src.game_environment import Game        # the reference class; __init__ needs the .npz, so it is bypassed
    imgs = F.make_sprites(SIZES, 0)
    s, r = _games(12, 1)
    big = imgs
    assert sum((m - 1) * b for m, b in zip(SIZES, F.REFERENCE_BASES)) < len(imgs)     # the shipped s_bases stay inside this table
    g = object.__new__(Game)
    g.games_no = 12
    g.imgs = torch.from_numpy(big.reshape(-1, 64, 64, 1))
    g.current_s = torch.from_numpy(s.copy())
    g.last_r = torch.from_numpy(r.copy())
    g.s_bases = torch.tensor(F.REFERENCE_BASES)
    ref = g.current_frame_all().numpy()                                  # (G,64,64,1)
    mine = F.current_frame_all(big, s, r, F.REFERENCE_BASES)             # (G,1,64,64)
    assert np.array_equal(ref.reshape(12, 4096), mine.reshape(12, 4096))
    assert int(g.s_to_index(g.current_s[3, :-1])) == F.s_to_index(s[3], F.REFERENCE_BASES)
    g.last_r[0] = 1.5
    with pytest.raises(ValueError):
        g.s_to_o(0)
    with pytest.raises(ValueError):
        F.s_to_o(big, s[0], 1.5, F.REFERENCE_BASES)


def test_oracle_against_golden_fixture():
    d = np.load(GOLD)
    imgs = F.make_sprites(SIZES, 0)
    assert np.array_equal(np.packbits(imgs.reshape(-1)), d["sprites_packed"])
    s, r = d["current_s"], d["last_r"]
    assert np.array_equal(F.current_frame_all(imgs, s, r, F.place_values(SIZES)), d["frames_place"])
    assert np.array_equal(F.current_frame_all(imgs, s, r, F.REFERENCE_BASES), d["frames_reference"])


def test_place_values_are_the_dsprites_ordering():
    assert F.place_values((1, 3, 6, 40, 32, 32)) == [737280, 245760, 40960, 1024, 32, 1]
    imgs = F.make_sprites(SIZES, 0)
    assert imgs.shape == (1536, 64, 64) and set(np.unique(imgs)) <= {0, 1}


@pytest.mark.gpu
def test_cuda_frames_equal_the_oracle_bit_for_bit():
    from dai_b200.game_environment import FrameProducer
    imgs = F.make_sprites(SIZES, 0)
    s, r = _games(300, 7)
    for bases, b in (("place", F.place_values(SIZES)), ("reference", F.REFERENCE_BASES)):
        fp = FrameProducer(imgs, SIZES, device="cuda:0", bases=bases)
        got = fp.current_frame_all(s, r)
        assert got.shape == (300, 1, 64, 64) and got.is_cuda
        assert np.array_equal(got.cpu().numpy(), F.current_frame_all(imgs, s, r, b))
        assert torch.equal(fp.s_to_index(torch.from_numpy(s)), torch.tensor([F.s_to_index(x, b) for x in s]))
        one = fp.current_frame(s, r, 5)
        assert one.shape == (64, 64, 1) and np.array_equal(one.cpu().numpy().reshape(64, 64), F.s_to_o(imgs, s[5], r[5], b))
    d = np.load(GOLD)
    fp = FrameProducer(imgs, SIZES, device="cuda:0")
    assert np.array_equal(fp.current_frame_all(d["current_s"], d["last_r"]).cpu().numpy(), d["frames_place"])


@pytest.mark.gpu
def test_cuda_frames_raise_where_the_reference_raises_and_feed_the_rollout():
    from dai_b200.game_environment import FrameProducer
    from dai_b200.torchmodel import ActiveInferenceModel
    import cases
    imgs = F.make_sprites(SIZES, 0)
    model = ActiveInferenceModel(10, 4, 1.0, 1.0, 1.0, device="cuda:0").load_numpy_weights(cases.weights_for("w0"))
    fp = FrameProducer(imgs, SIZES, engine=model._engine)
    s, r = _games(6, 3)
    bad_r = r.copy(); bad_r[4] = 1.25
    with pytest.raises(ValueError):
        fp.current_frame_all(s, bad_r)
    bad_s = s.copy(); bad_s[2, 1] = 99
    with pytest.raises(ValueError):
        fp.current_frame_all(bad_s, r)
    # frames born on the device go straight into the path: same G as the same frames passed from the host
    o = fp.current_frame_all(s, r)
    model.set_rng(5, 0)
    G1, _, _ = model.calculate_G_repeated(o.repeat_interleave(4, dim=0), torch.eye(4).repeat(6, 1), steps=1, samples=2)
    model.set_rng(5, 0)
    host = torch.from_numpy(F.current_frame_all(imgs, s, r, F.place_values(SIZES)))
    G2, _, _ = model.calculate_G_repeated(host.repeat_interleave(4, dim=0), torch.eye(4).repeat(6, 1), steps=1, samples=2)
    assert torch.equal(G1, G2)

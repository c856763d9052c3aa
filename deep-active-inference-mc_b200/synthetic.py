"""Synthetic weights and dSprites-like frames for the EFE-rollout path.

There is no trained checkpoint and no dSprites .npz in the reference snapshot
(SURVEY.md §4), so every parity test and benchmark runs on the generators
below.  They are numpy-only (PCG64, stable across numpy versions) so that the
CPU box that writes the golden fixtures and the GPU box that checks them build
bit-identical tensors.

Shapes follow the reference's state_dict (src/torchmodel.py:19-25, 41-52,
84-128) with the D1 repair (encoder FC1 is 576->256).
"""
import numpy as np

S_DIM = 10
PI_DIM = 4
RES = 64

# (key, shape, fan_in).  Conv weights are torch layout: Conv2d (Cout,Cin,kh,kw),
# ConvTranspose2d (Cin,Cout,kh,kw).  fan_in mirrors torch's default init, which
# for ConvTranspose2d is size(1)*kh*kw (src/torchmodel.py relies on nn defaults).
WEIGHT_SPECS = [
    # model_top (src/torchmodel.py:19-25)
    ("qpi_net.0", (128, 10), 10),
    ("qpi_net.2", (128, 128), 128),
    ("qpi_net.4", (4, 128), 128),
    # model_mid (src/torchmodel.py:41-52)
    ("ps_net.0", (512, 14), 14),
    ("ps_net.3", (512, 512), 512),
    ("ps_net.6", (512, 512), 512),
    ("ps_net.9", (20, 512), 512),
    # model_down.qs_net (src/torchmodel.py:84-104)
    ("qs_net.0", (32, 1, 3, 3), 9),
    ("qs_net.2", (32, 32, 3, 3), 288),
    ("qs_net.4", (64, 32, 3, 3), 288),
    ("qs_net.6", (64, 64, 3, 3), 576),
    ("qs_net.9", (256, 576), 576),
    ("qs_net.12", (256, 256), 256),
    ("qs_net.15", (256, 256), 256),
    ("qs_net.18", (20, 256), 256),
    # model_down.po_net (src/torchmodel.py:106-128)
    ("po_net.0", (256, 10), 10),
    ("po_net.3", (256, 256), 256),
    ("po_net.6", (256, 256), 256),
    ("po_net.9", (16384, 256), 256),
    ("po_net.13", (64, 64, 3, 3), 576),
    ("po_net.15", (64, 64, 3, 3), 576),
    ("po_net.17", (64, 32, 3, 3), 288),
    ("po_net.19", (32, 1, 3, 3), 9),
]

MODULE_OF = {"qpi_net": "model_top", "ps_net": "model_mid", "qs_net": "model_down", "po_net": "model_down"}


def bias_shape(key, wshape):
    """Bias length: Cout.  ConvTranspose2d stores (Cin,Cout,..) so Cout = shape[1]."""
    if key.startswith("po_net.1") and len(wshape) == 4:
        return (wshape[1],)
    return (wshape[0],)


def make_weights(seed=0, sharp=False):
    """All 46 state_dict tensors as float32 numpy arrays, keyed '<net>.<idx>.weight|bias'.

    U(-1/sqrt(fan_in), 1/sqrt(fan_in)) like torch's default Linear/Conv init.
    sharp=True scales the last deconv weight x40 so the sigmoid saturates
    (SURVEY.md Appendix B), which is the hard case for the 1e-4 tolerance.
    """
    rng = np.random.default_rng(seed)
    out = {}
    for key, shape, fan_in in WEIGHT_SPECS:
        bound = 1.0 / np.sqrt(fan_in)
        out[key + ".weight"] = rng.uniform(-bound, bound, size=shape).astype(np.float32)
        out[key + ".bias"] = rng.uniform(-bound, bound, size=bias_shape(key, shape)).astype(np.float32)
    if sharp:
        out["po_net.19.weight"] = (out["po_net.19.weight"] * np.float32(40.0)).astype(np.float32)
    return out


def make_frames(n, seed=0):
    """n dSprites-like observations, float32 NCHW (n,1,64,64).

    A binary blob (ellipse / square / heart-ish) of random scale, orientation
    and position on rows 3..63, plus the reward bar the environment paints on
    rows 0..2 (src/game_environment.py:44-54): columns 0..31 = r for r >= 0,
    columns 32..63 = -r for r < 0, r ~ U(-1,1).
    """
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:RES, 0:RES].astype(np.float32)
    frames = np.zeros((n, 1, RES, RES), dtype=np.float32)
    for i in range(n):
        shape = rng.integers(0, 3)
        scale = rng.uniform(3.0, 8.0)
        theta = rng.uniform(0.0, 2.0 * np.pi)
        cx = rng.uniform(10.0, 54.0)
        cy = rng.uniform(13.0, 54.0)
        c, s = np.cos(theta), np.sin(theta)
        u = ((xx - cx) * c + (yy - cy) * s) / scale
        v = (-(xx - cx) * s + (yy - cy) * c) / scale
        if shape == 0:
            blob = (u * u + (v * v) * 2.25) <= 1.0
        elif shape == 1:
            blob = (np.abs(u) <= 0.8) & (np.abs(v) <= 0.8)
        else:
            blob = (u * u + (1.2 * v - np.sqrt(np.abs(u))) ** 2) <= 1.0
        img = blob.astype(np.float32)
        img[0:3, :] = 0.0
        r = rng.uniform(-1.0, 1.0)
        if r >= 0.0:
            img[0:3, 0:32] = r
        else:
            img[0:3, 32:64] = -r
        frames[i, 0] = img
    return frames

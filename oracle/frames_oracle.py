"""Oracle (test infrastructure, CPU, numpy) for the frame producer in front of the EFE path — a restatement of
Game.s_to_index / s_to_o / current_frame_all of the reference's environment (src/game_environment.py:39-66).

Only tests/, __graft_entry__.smoke() and bench.py's cpu legs may import this module; the product path is the CUDA
kernel behind dai_frames_render.  Pinned against the reference's own Game methods (run on a synthetic sprite table,
tests/test_frames.py) where /root/reference is mounted, and through tests/golden/frames_golden.npz elsewhere.
"""
import numpy as np

REFERENCE_BASES = (1, 3, 6, 40, 32, 32)      # Game.s_bases as shipped (src/game_environment.py:24-25)


def place_values(latents_sizes):
    """Mixed-radix place values of the dSprites latent classes: index = sum_i class_i * place_i (the dataset's own
    ordering; for [1,3,6,40,32,32]: [737280, 245760, 40960, 1024, 32, 1])."""
    sizes = [int(x) for x in latents_sizes]
    out, pv = [0] * len(sizes), 1
    for i in range(len(sizes) - 1, -1, -1):
        out[i] = pv
        pv *= sizes[i]
    return out


def s_to_index(s, bases):
    """src/game_environment.py:39-42: s.to(int64) (truncation) dotted with the bases."""
    return int(np.dot(np.trunc(np.asarray(s[:6], dtype=np.float64)).astype(np.int64), np.asarray(bases, dtype=np.int64)))


def s_to_o(imgs, s, r, bases):
    """src/game_environment.py:44-54.  imgs (count,64,64) uint8 -> (64,64) float32 with the reward bar."""
    idx = s_to_index(s, bases)
    if idx < 0 or idx >= imgs.shape[0]:
        raise IndexError("sprite index %d outside the table" % idx)
    o = (imgs[idx].reshape(64, 64) != 0).astype(np.float32)
    r = np.float32(r)
    if 0.0 <= r <= 1.0:
        o[0:3, 0:32] = r
    elif -1.0 <= r < 0.0:
        o[0:3, 32:64] = -r
    else:
        raise ValueError("Error: Reward: %r" % (r,))
    return o


def current_frame_all(imgs, current_s, last_r, bases):
    """src/game_environment.py:62-66 -> (G,1,64,64) float32 (the model's NCHW view of the (G,64,64,1) frames)."""
    return np.stack([s_to_o(imgs, current_s[i], last_r[i], bases) for i in range(len(last_r))])[:, None]


def make_sprites(latents_sizes, seed=0):
    """A synthetic binary sprite table with dSprites' latent structure (colour, shape, scale, orientation, x, y) at any
    `latents_sizes` — the real .npz is not in the snapshot.  (count,64,64) uint8."""
    sizes = [int(x) for x in latents_sizes]
    count = int(np.prod(sizes))
    yy, xx = np.mgrid[0:64, 0:64].astype(np.float32)
    imgs = np.zeros((count, 64, 64), dtype=np.uint8)
    pv = place_values(sizes)
    for idx in range(count):
        c = [(idx // pv[i]) % sizes[i] for i in range(6)]
        shape, scale = c[1] % 3, 3.0 + 5.0 * (c[2] + 1) / sizes[2]
        th = 2.0 * np.pi * c[3] / sizes[3]
        cx, cy = 8.0 + 48.0 * c[4] / max(sizes[4] - 1, 1), 8.0 + 48.0 * c[5] / max(sizes[5] - 1, 1)
        u = ((xx - cx) * np.cos(th) + (yy - cy) * np.sin(th)) / scale
        v = (-(xx - cx) * np.sin(th) + (yy - cy) * np.cos(th)) / scale
        if shape == 0:
            blob = (np.abs(u) <= 0.8) & (np.abs(v) <= 0.8)
        elif shape == 1:
            blob = (u * u + 2.25 * v * v) <= 1.0
        else:
            blob = (u * u + (1.2 * v - np.sqrt(np.abs(u))) ** 2) <= 1.0
        imgs[idx] = blob
    return imgs

"""Writes tests/golden/frames_golden.npz: frames produced by the REFERENCE's own Game.current_frame_all
(src/game_environment.py:62-66, imported read-only from /root/reference; __init__ bypassed because the dSprites .npz is
not in the snapshot) on the synthetic sprite table of oracle/frames_oracle.make_sprites, for the reference's s_bases,
and by the oracle for the place-value bases (the reference has no such mode).  Run in the authoring container:
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_frames_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
from oracle import frames_oracle as F          # noqa: E402
from src.game_environment import Game          # noqa: E402

SIZES = (1, 3, 2, 4, 8, 8)
imgs = F.make_sprites(SIZES, 0)
rng = np.random.default_rng(11)
n = 24
s = np.zeros((n, 7), dtype=np.float32)
for i, m in enumerate(SIZES):
    s[:, i] = rng.integers(0, m, size=n)
r = rng.uniform(-1, 1, size=n).astype(np.float32)
r[0], r[1], r[2] = 0.0, -1.0, 1.0

g = object.__new__(Game)
g.games_no = n
g.imgs = torch.from_numpy(imgs.reshape(-1, 64, 64, 1))
g.current_s = torch.from_numpy(s.copy())
g.last_r = torch.from_numpy(r.copy())
g.s_bases = torch.tensor(F.REFERENCE_BASES)
ref = g.current_frame_all().numpy().reshape(n, 1, 64, 64)
assert np.array_equal(ref, F.current_frame_all(imgs, s, r, F.REFERENCE_BASES)), "oracle != reference"
g.s_bases = torch.tensor(F.place_values(SIZES))      # the reference's own code with the repaired bases
place = g.current_frame_all().numpy().reshape(n, 1, 64, 64)
assert np.array_equal(place, F.current_frame_all(imgs, s, r, F.place_values(SIZES))), "oracle != reference (place values)"
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "frames_golden.npz"), sprites_packed=np.packbits(imgs.reshape(-1)),
                    current_s=s, last_r=r, frames_reference=ref, frames_place=place)
print("wrote frames_golden.npz", ref.shape)

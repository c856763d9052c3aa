"""Device-side frame producer — the step in front of the EFE path.

Mirrors the frame-producing methods of the reference's environment, `Game.s_to_index`, `Game.s_to_o`,
`Game.current_frame`, `Game.current_frame_all` (src/game_environment.py:39-66): the latent classes `current_s` and the
reward `last_r` of all games go to ONE kernel (dai_frames_render) that gathers the sprites from a bit-packed table in
HBM and paints the reward bar, instead of a Python loop over games; the frames are born on the device, NCHW, ready for
`calculate_G_repeated`.  The game rules (moves, rewards, sampling) are control plane and stay with the caller.

`bases`: "place" (default) = the mixed-radix place values of latents_sizes, i.e. the dataset's real indexing;
"reference" = Game.s_bases as shipped ([1,3,6,40,32,32]), for bit-parity with the reference.
"""
import numpy as np
import torch

from .engine import Engine

DSPRITES_FILE = "dsprites_ndarray_co1sh3sc6or40x32y32_64x64.npz"      # src/game_environment.py:10
DSPRITES_SIZES = (1, 3, 6, 40, 32, 32)


class FrameProducer:
    def __init__(self, imgs, latents_sizes=DSPRITES_SIZES, engine=None, device=None, bases="place"):
        if bases not in ("place", "reference"):
            raise ValueError("bases must be 'place' or 'reference'")
        self.engine = engine if engine is not None else Engine(device=device)
        self.latents_sizes = tuple(int(x) for x in latents_sizes)
        self.reference_bases = bases == "reference"
        pv, out = 1, []
        for n in reversed(self.latents_sizes):
            out.append(pv)
            pv *= n
        self.s_bases = torch.tensor([1, 3, 6, 40, 32, 32] if self.reference_bases else out[::-1])
        self.engine.set_sprites(imgs, self.latents_sizes)

    @classmethod
    def from_npz(cls, path=DSPRITES_FILE, **kw):
        """The reference's dataset file (src/game_environment.py:10-16)."""
        d = np.load(path, allow_pickle=True, encoding="latin1")
        sizes = d["metadata"][()]["latents_sizes"]
        return cls(d["imgs"], sizes, **kw)

    def s_to_index(self, s):
        """src/game_environment.py:39-42."""
        s = torch.as_tensor(s)[..., :6].to(dtype=self.s_bases.dtype)
        return (s * self.s_bases).sum(-1).long()

    def current_frame_all(self, current_s, last_r):
        """(G, >=6) latent classes, (G,) rewards -> (G,1,64,64) float32 CUDA tensor (src/game_environment.py:62-66;
        the reference returns (G,64,64,1) — the same memory, C = 1).  Raises ValueError where the reference raises."""
        return self.engine.render_frames(current_s, last_r, reference_bases=self.reference_bases)

    def current_frame(self, current_s, last_r, index):
        """src/game_environment.py:59-60 -> (64,64,1)."""
        s = torch.as_tensor(current_s)[index:index + 1]
        r = torch.as_tensor(last_r).reshape(-1)[index:index + 1]
        return self.current_frame_all(s, r).reshape(64, 64, 1)

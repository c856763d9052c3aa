"""Generate tests/golden/efe_golden.npz from the REAL reference (authoring container only).

    cd /root/repo && python tests/golden/make_golden.py

For every case in tests/cases.py:
  1. run the oracle (oracle/efe_oracle.py) with keyed Philox noise, recording every draw
     on a tape in the reference's RNG order;
  2. import the reference from /root/reference (read-only, never copied), apply the two
     runtime shims of SURVEY.md §0.1 (D1: encoder FC1 576->256, D2: `precision`), load
     the same synthetic weights, patch torch.nn.functional.dropout / torch.randn_like /
     torch.multinomial to replay that tape, and run the same case on it;
  3. require oracle == reference bit for bit (same ATen ops, same inputs, same noise),
     and that the tape was consumed exactly;
  4. store the REFERENCE's outputs.
It also checks the oracle bit-for-bit against the reference under torch's own RNG stream
(TorchStreamNoise) for every evaluator.  The .npz is what pins parity on the GPU box,
where /root/reference does not exist.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")

import dai_b200  # noqa: E402,F401
import cases  # noqa: E402
from oracle import efe_oracle as O  # noqa: E402
from src.torchmodel import ActiveInferenceModel  # noqa: E402  (the reference)


def make_reference(weights):
    m = ActiveInferenceModel(10, 4, 1.0, 1.0, 1.0)
    m.model_down.qs_net[9] = torch.nn.Linear(576, 256)      # SHIM-1 (D1)
    m.precision = torch.float32                             # SHIM-2 (D2)
    for mod in (m.model_top, m.model_mid, m.model_down):
        mod.load_state_dict({k: torch.from_numpy(weights[k]) for k in mod.state_dict()})
    m.set_rng = lambda seed, call=0: None
    return m


class Replay:
    """Serve the oracle's recorded draws to the reference, in order."""

    def __init__(self, tape):
        self.tape, self.pos = tape, 0

    def _next(self, kind):
        k, v = self.tape[self.pos]
        assert k == kind, "draw %d: reference wants %s, tape has %s" % (self.pos, kind, k)
        self.pos += 1
        return v

    def __enter__(self):
        import torch.nn.functional as F
        self.saved = (F.dropout, torch.randn_like, torch.multinomial)

        def dropout(x, p=0.5, training=True, inplace=False):
            m = self._next("d")
            assert m.shape == x.shape, (m.shape, x.shape)
            return x * m

        def randn_like(x, *a, **k):
            e = self._next("r")
            assert e.shape == x.shape, (e.shape, x.shape)
            return e

        def multinomial(q, n, *a, **k):
            return torch.tensor([self._next("m")])

        F.dropout, torch.randn_like, torch.multinomial = dropout, randn_like, multinomial
        return self

    def __exit__(self, *exc):
        import torch.nn.functional as F
        F.dropout, torch.randn_like, torch.multinomial = self.saved


def stream_pin(ref, W):
    """Oracle == reference bit for bit under torch's global RNG stream."""
    import dai_b200.synthetic as syn
    o = torch.from_numpy(syn.make_frames(1, 3)).repeat(4, 1, 1, 1)
    eye = torch.eye(4)

    def both(f_ref, f_ora, seed):
        torch.manual_seed(seed)
        a = f_ref()
        torch.manual_seed(seed)
        b = f_ora()
        return a, b

    def flat(x):
        out = []
        for v in (x if isinstance(x, (tuple, list)) else [x]):
            if isinstance(v, (tuple, list)):
                out += flat(v)
            elif isinstance(v, torch.Tensor):
                out.append(v.detach())
            else:
                out.append(torch.tensor(v))
        return out

    s0 = torch.randn(4, 10, generator=torch.Generator().manual_seed(1))
    checks = [
        ("calculate_G", lambda: ref.calculate_G(s0, eye, samples=3),
         lambda: O.calculate_G(W, s0, eye, 3, O.TorchStreamNoise())),
        ("calculate_G_mean", lambda: ref.calculate_G_mean(s0, eye),
         lambda: O.calculate_G_mean(W, s0, eye, O.TorchStreamNoise())),
        ("calculate_G_4_repeated", lambda: ref.calculate_G_4_repeated(o, steps=2, samples=2),
         lambda: O.calculate_G_repeated(W, o, None, 2, False, 2, O.TorchStreamNoise(), four=True)),
        ("calculate_G_4_repeated(mean)", lambda: ref.calculate_G_4_repeated(o, steps=2, calc_mean=True),
         lambda: O.calculate_G_repeated(W, o, None, 2, True, 10, O.TorchStreamNoise(), four=True)),
        ("calculate_G_repeated", lambda: ref.calculate_G_repeated(o, eye, steps=2, samples=2, calc_mean=True),
         lambda: O.calculate_G_repeated(W, o, eye, 2, True, 2, O.TorchStreamNoise())),
        ("mcts_step_simulate", lambda: ref.mcts_step_simulate(s0[0], 3),
         lambda: O.mcts_step_simulate(W, s0[0], 3, False, O.TorchStreamNoise(), O.TorchStreamNoise())),
        ("trajectory", lambda: ref.calculate_G_given_trajectory(s0, s0 * 0.5, s0 * 0.3, s0 * 0.1, eye),
         lambda: O.calculate_G_given_trajectory(W, s0, s0 * 0.5, s0 * 0.3, s0 * 0.1, eye, O.TorchStreamNoise())),
    ]
    for i, (name, fr, fo) in enumerate(checks):
        a, b = both(fr, fo, 100 + i)
        for x, y in zip(flat(a), flat(b)):
            assert torch.equal(x, y), "stream pin failed: " + name
        print("stream pin ok (bit-exact):", name)


def main():
    torch.set_num_threads(8)
    out = {}
    refs, oras = {}, {}
    for name, (wk, _) in cases.CASES.items():
        if wk not in refs:
            w = cases.weights_for(wk)
            refs[wk] = make_reference(w)
            oras[wk] = O.OracleModel(w, seed=cases.SEED)
            if wk == "w0":
                stream_pin(refs[wk], oras[wk].W)
        ora, ref = oras[wk], refs[wk]
        ora.tape = []
        got_o = cases.run_case(name, ora)
        tape, ora.tape = ora.tape, None
        with Replay(tape) as rp:
            got_r = cases.run_case(name, ref)
        assert rp.pos == len(tape), "%s: tape %d of %d consumed" % (name, rp.pos, len(tape))
        for k in got_r:
            assert np.array_equal(got_o[k], got_r[k]), "%s.%s oracle != reference" % (name, k)
            out[name + "/" + k] = got_r[k]
        print("replay pin ok (bit-exact): %-28s draws=%d fields=%s" % (name, len(tape), ",".join(got_r)))
    path = os.path.join(ROOT, "tests", "golden", "efe_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

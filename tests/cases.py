"""Parity cases for the EFE-rollout path, written against the reference's call surface
(src/torchmodel.py:149-393).  The same `run_case` drives the oracle (CPU), the real
reference under replayed noise (tests/golden/make_golden.py, authoring container only)
and the CUDA model (GPU tests), so the three are compared on identical inputs.

Inputs are rebuilt from seeds (nothing big is stored); outputs are flattened to
{name: float32 ndarray}.  Noise is keyed: every case starts at (SEED, call index).
"""
import numpy as np
import torch

import dai_b200.synthetic as syn

SEED = 1234

# name -> (weights kind, callable(model, dev) -> dict of tensors)
CASES = {}


def case(name, weights=("w0",)):
    def deco(fn):
        for w in weights:
            CASES[name + "@" + w] = (w, fn)
        return fn
    return deco


def _rand(shape, seed, scale=1.0):
    return torch.from_numpy((np.random.default_rng(seed).standard_normal(shape) * scale).astype(np.float32))


def _eye(dev, reps=1):
    return torch.eye(4, device=dev).repeat(reps, 1)


def weights_for(kind):
    return syn.make_weights(0, sharp=(kind == "w0s"))


@case("encoder", ("w0",))
def _encoder(m, dev):
    o = torch.from_numpy(syn.make_frames(3, 1)).to(dev)
    mean, logvar = m.model_down.encoder(o)
    s, mean2, logvar2 = m.model_down.encoder_with_sample(o)
    return dict(mean=mean, logvar=logvar, s=s, mean2=mean2, logvar2=logvar2)


@case("transition", ("w0",))
def _transition(m, dev):
    s0 = _rand((4, 10), 2).to(dev)
    mean, logvar = m.model_mid.transition(_eye(dev), s0)
    ps1, mean2, logvar2 = m.model_mid.transition_with_sample(_eye(dev), s0)
    return dict(mean=mean, logvar=logvar, ps1=ps1, mean2=mean2, logvar2=logvar2)


@case("decoder", ("w0", "w0s"))
def _decoder(m, dev):
    s = _rand((4, 10), 3).to(dev)
    return dict(po=m.model_down.decoder(s))


@case("habit", ("w0",))
def _habit(m, dev):
    s = _rand((5, 10), 4).to(dev)
    logits, q, logq = m.model_top.encode_s(s)
    return dict(logits=logits, q=q, logq=logq)


@case("reward", ("w0",))
def _reward(m, dev):
    o = torch.from_numpy(syn.make_frames(3, 5)).to(dev)
    return dict(r=m.check_reward(o))


@case("calculate_G", ("w0", "w0s"))
def _calc_g(m, dev):
    s0 = _rand((4, 10), 6).to(dev)
    G, terms, ps1, ps1_mean, po1 = m.calculate_G(s0, _eye(dev), samples=3)
    return dict(G=G, t0=terms[0], t1=terms[1], t2=terms[2], ps1=ps1, ps1_mean=ps1_mean, po1=po1)


@case("calculate_G_mean", ("w0", "w0s"))
def _calc_g_mean(m, dev):
    s0 = _rand((4, 10), 7).to(dev)
    G, terms, ps1_mean, po1 = m.calculate_G_mean(s0, _eye(dev))
    return dict(G=G, t0=terms[0], t1=terms[1], t2=terms[2], ps1_mean=ps1_mean, po1=po1)


def _g4(m, dev, steps, samples, calc_mean, frame_seed):
    o = torch.from_numpy(syn.make_frames(1, frame_seed)).to(dev).repeat(4, 1, 1, 1)   # test_demo.py:149
    G, terms, po1 = m.calculate_G_4_repeated(o, steps=steps, calc_mean=calc_mean, samples=samples)
    return dict(G=G, t0=terms[0], t1=terms[1], t2=terms[2], po1=po1)


@case("g4_repeated", ("w0", "w0s"))
def _g4_rep(m, dev):
    return _g4(m, dev, 2, 2, False, 8)


@case("g4_repeated_mean", ("w0",))
def _g4_rep_mean(m, dev):
    return _g4(m, dev, 3, 1, True, 9)


@case("config1", ("w0", "w0s"))
def _config1(m, dev):
    """BASELINE.json configs[0]: N=10, T=1 (test_demo.py:150)."""
    return _g4(m, dev, 1, 10, False, 10)


def _grep(m, dev, calc_mean):
    o = torch.from_numpy(syn.make_frames(2, 11)).to(dev).repeat_interleave(4, dim=0)   # util.py:57
    G, terms, po1 = m.calculate_G_repeated(o, _eye(dev, 2), steps=2, calc_mean=calc_mean, samples=2)
    return dict(G=G, t0=terms[0], t1=terms[1], t2=terms[2], po1=po1)


@case("g_repeated", ("w0",))
def _g_rep(m, dev):
    return _grep(m, dev, False)


@case("g_repeated_mean", ("w0",))
def _g_rep_mean(m, dev):
    return _grep(m, dev, True)


@case("trajectory", ("w0", "w0s"))
def _traj(m, dev):
    d = 3
    s0, ps1, mu = _rand((d, 10), 12).to(dev), _rand((d, 10), 13).to(dev), _rand((d, 10), 14).to(dev)
    lv = _rand((d, 10), 15, 0.3).to(dev)
    pi = torch.eye(4, device=dev)[[2, 0, 3]]
    return dict(G=m.calculate_G_given_trajectory(s0, ps1, mu, lv, pi))


@case("simulate", ("w0",))
def _sim(m, dev):
    out = {}
    for i, (depth, use_means) in enumerate(((3, False), (1, False), (4, True))):
        G, pi0, qpi = m.mcts_step_simulate(_rand((10,), 16 + i).to(dev), depth, use_means=use_means)
        out.update({"G%d" % i: torch.tensor([G]), "pi0_%d" % i: pi0, "qpi%d" % i: qpi})
    return out


@case("imagine_habit", ("w0",))
def _imagine(m, dev):
    o = torch.from_numpy(syn.make_frames(2, 20)).to(dev)
    pi = torch.eye(4, device=dev)[[1, 3]]
    return dict(po=m.imagine_future_from_o(o, pi), qpi=m.habitual_net(o))


def run_case(name, model, dev="cpu", call=0):
    """Run one case at (SEED, call); returns {field: float32 ndarray}."""
    _, fn = CASES[name]
    model.set_rng(SEED, call)
    with torch.no_grad():
        out = fn(model, dev)
    return {k: v.detach().to("cpu", torch.float32).numpy() for k, v in out.items()}


# Parity definition: relative 1e-4 for G, term0, term1; term2 is a
# cancelling difference so it is judged against |G|; images / latents rtol 1e-4 + atol 1e-5.
RTOL = 1e-4
ATOL = 1e-5


def compare(name, got, ref, rtol=RTOL, atol=ATOL):
    """Returns a list of human-readable mismatches (empty = parity)."""
    bad = []
    for k in ref:
        a, b = np.asarray(got[k], dtype=np.float64), np.asarray(ref[k], dtype=np.float64)
        if a.shape != b.shape:
            bad.append("%s.%s shape %s vs %s" % (name, k, a.shape, b.shape))
            continue
        if k.startswith("pi0"):
            if not np.array_equal(a, b):
                bad.append("%s.%s action sequence differs" % (name, k))
            continue
        if k == "t2":
            scale = np.abs(np.asarray(ref["G"], dtype=np.float64))
            err = np.abs(a - b) - rtol * scale
        elif k in ("G", "t0", "t1", "r") or k.startswith("G"):
            err = np.abs(a - b) - rtol * np.abs(b)
        else:
            err = np.abs(a - b) - (atol + rtol * np.abs(b))
        if not np.all(np.isfinite(a)) or err.max() > 0:
            i = int(np.argmax(err))
            bad.append("%s.%s max excess %.3e at %d (got %.7g ref %.7g)" %
                       (name, k, float(err.max()), i, a.flat[i], b.flat[i]))
    return bad


class TeacherForced:
    """Model stand-in for planner tests: every call the planner makes is served by the CUDA model AND evaluated by
    the oracle on the SAME inputs under the SAME noise key (the oracle's call index is set to the engine's before each
    call); the two answers must agree to the parity tolerance, and the planner continues on the CUDA answer.

    Why not two independent searches: a search is a chain of argmax decisions over Q values that differ between actions
    by ~1e-3 of |G| on random-init nets, so a 1e-6 relative difference in one G (fp32 CPU vs bf16x3 tensor cores, both
    inside the 1e-4 bar) can flip a near-tie after a few dozen expansions and the two trees then differ legitimately.
    Teacher forcing checks what parity means for a planner: along the search the CUDA model actually drives, each
    evaluation equals the reference arithmetic."""

    def __init__(self, gpu, ora, seed):
        self.gpu, self.ora, self.seed = gpu, ora, seed
        self.pi_dim, self.s_dim = gpu.pi_dim, gpu.s_dim
        self.pi_one_hot, self.pi_one_hot_3 = gpu.pi_one_hot, gpu.pi_one_hot_3
        self.device = gpu.device
        self.calls = 0
        self.worst = {}
        outer = self

        class _Down:
            resolution = 64

            def encoder(self, o):
                return outer._both("encoder", lambda m, x: m.model_down.encoder(x), o)

        class _Top:
            def encode_s(self, s):
                return outer._both("encode_s", lambda m, x: m.model_top.encode_s(x), s)

        self.model_down, self.model_top = _Down(), _Top()

    def _cpu(self, x):
        return x.detach().cpu() if isinstance(x, torch.Tensor) else x

    def _flat(self, out):
        res = []
        for v in (out if isinstance(out, (tuple, list)) else [out]):
            if isinstance(v, (tuple, list)):
                res += self._flat(v)
            elif isinstance(v, torch.Tensor):
                res.append(v.detach().to("cpu", torch.float64).reshape(-1).numpy())
            else:
                res.append(np.asarray([float(v)], dtype=np.float64))
        return res

    def _both(self, name, fn, *args, exact=(), scalar_rel=()):
        self.ora.set_rng(self.seed, self.gpu._engine.get_rng()[1])
        got = fn(self.gpu, *args)
        with torch.no_grad():
            ref = fn(self.ora, *[self._cpu(a) for a in args])
        assert self.gpu._engine.get_rng()[1] == self.ora.call, name
        for i, (a, b) in enumerate(zip(self._flat(got), self._flat(ref))):
            assert a.shape == b.shape, (name, i)
            if i in exact:
                assert np.array_equal(a, b), "%s output %d differs (call %d)" % (name, i, self.calls)
                continue
            tol = RTOL * np.abs(b) + (0.0 if i in scalar_rel else ATOL)
            err = np.abs(a - b)
            assert np.all(np.isfinite(a)) and np.all(err <= tol), \
                "%s output %d: max excess %.3e (call %d)" % (name, i, float((err - tol).max()), self.calls)
            self.worst[name] = max(self.worst.get(name, 0.0), float((err / (np.abs(b) + 1e-30)).max()) if i in scalar_rel else 0.0)
        self.calls += 1
        return got

    def calculate_G(self, s0, pi0, samples=10):
        # G relative 1e-4; terms: t0, t1 relative, t2 against |G| -> checked through G (= -t0 + t1 + t2); latents/images rtol+atol
        out = self._both("calculate_G", lambda m, s, p: (lambda r: (r[0], r[2], r[3], r[4]))(m.calculate_G(s, p, samples=samples)),
                         s0, pi0, scalar_rel=(0,))
        return out[0], None, out[1], out[2], out[3]

    def calculate_G_mean(self, s0, pi0):
        out = self._both("calculate_G_mean", lambda m, s, p: (lambda r: (r[0], r[2], r[3]))(m.calculate_G_mean(s, p)), s0, pi0,
                         scalar_rel=(0,))
        return out[0], None, out[1], out[2]

    def mcts_step_simulate(self, starting_s, depth, use_means=False):
        return self._both("mcts_step_simulate", lambda m, s: m.mcts_step_simulate(s, depth, use_means=use_means), starting_s,
                          exact=(1,), scalar_rel=(0,))

"""TEST INFRASTRUCTURE — not part of the product path.

CPU (torch fp32) restatement of the reference's Monte-Carlo EFE rollout path:
src/torchmodel.py:10-146 (nets), :210-393 (EFE evaluators) and
src/torchutils.py:19-37 (entropies, preferred-outcome reward) of
zfountas/deep-active-inference-mc @ d7e76d8, with the D1 repair (encoder FC1 is
576->256).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` legs may import it; the product path
(deep-active-inference-mc_b200/) never does.

Parity pin: the reference ships no golden vectors or tests, so
the oracle is pinned against the reference itself executed in the authoring
container — (1) bit-for-bit under the same torch seed (TorchStreamNoise, the
oracle consumes torch's global generator in exactly the reference's order) and
(2) under keyed Philox noise replayed into the reference through patched
F.dropout / torch.randn_like / torch.multinomial.  tests/golden/make_golden.py
does both and writes the fixtures tests/golden/*.npz that travel to the GPU box
(where /root/reference does not exist).

Every function names the reference lines it follows.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import philox
from .philox import SITES

S_DIM = 10
PI_DIM = 4

# src/torchutils.py:16 — a float64 numpy scalar in the reference; results stay float32
LOG_2_PI_E = np.log(2.0 * np.pi * np.e)
DISPLACEMENT = 0.00001   # src/torchutils.py:26,30


# --------------------------------------------------------------------------- noise

class _Noise:
    """Noise source with a (step, sample) cursor.  training=False => dropout is identity
    (nn.Dropout in eval mode); the normals are still drawn (the reference never gates them)."""

    def __init__(self, training=True):
        self.training = training
        self.step = 0
        self.sample = 0

    def at(self, step=None, sample=None):
        if step is not None:
            self.step = int(step)
        if sample is not None:
            self.sample = int(sample)
        return self


class TorchStreamNoise(_Noise):
    """Draws from torch's global CPU generator through the same calls the reference makes
    (nn.Dropout -> F.dropout, torch.randn_like, torch.multinomial), so with the same
    torch.manual_seed the oracle reproduces the reference bit for bit.  Also the noise the
    CPU-baseline timing uses: its cost is the reference's cost (bernoulli_ = 23 % of wall)."""

    def dropout(self, x, site):
        return F.dropout(x, 0.5, self.training)

    def randn_like(self, x, site):
        return torch.randn_like(x)

    def categorical(self, q, site, row=0):
        return int(torch.multinomial(q, 1).item())


class PhiloxNoise(_Noise):
    """Keyed counter-based noise (oracle/philox.py), identical to what the CUDA kernels
    generate.  `tape`, when a list, records every draw in call order so it can be replayed
    into the real reference (tests/golden/make_golden.py)."""

    def __init__(self, key, training=True, tape=None):
        super().__init__(training)
        self.key = int(key)
        self.tape = tape
        self.row0 = 0          # noise row of batch row 0 (batched simulations: rollout k draws as row k)

    def dropout(self, x, site):
        if not self.training:
            return x
        m = torch.from_numpy(philox.dropout_mask(self.key, self.step, self.sample, site,
                                                 self.row0 + np.arange(x.shape[0]), x.shape[1]))
        if self.tape is not None:
            self.tape.append(("d", m))
        return x * m

    def randn_like(self, x, site):
        e = torch.from_numpy(philox.normals(self.key, self.step, self.sample, site,
                                            self.row0 + np.arange(x.shape[0]), x.shape[1]))
        if self.tape is not None:
            self.tape.append(("r", e))
        return e

    def categorical(self, q, site, row=0):
        """Inverse-CDF draw in float32: first i with u*total < cumsum_i (fallback last)."""
        u = philox.uniform24(self.key, self.step, self.sample, site, self.row0 + row)
        c = np.float32(0.0)
        cdf = []
        for v in q.detach().numpy().astype(np.float32):
            c = np.float32(c + v)
            cdf.append(c)
        thr = np.float32(u * cdf[-1])
        a = len(cdf) - 1
        for i, ci in enumerate(cdf):
            if thr < ci:
                a = i
                break
        if self.tape is not None:
            self.tape.append(("m", a))
        return a


# --------------------------------------------------------------------------- nets

def to_torch(weights):
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in weights.items()}


def _fc(W, name, x):
    return F.linear(x, W[name + ".weight"], W[name + ".bias"])


def qpi_forward(W, s):
    """ModelTop.encode_s — src/torchmodel.py:19-31."""
    h = F.relu(_fc(W, "qpi_net.0", s))
    h = F.relu(_fc(W, "qpi_net.2", h))
    logits = _fc(W, "qpi_net.4", h)
    q = F.softmax(logits, dim=-1)
    return logits, q, torch.log(q + 1e-20)


def ps_forward(W, pi, s0, nz, site):
    """ModelMid.transition — src/torchmodel.py:41-52,58-61.  pi is concatenated first."""
    h = torch.cat([pi, s0], dim=1)
    for j, idx in enumerate((0, 3, 6)):
        h = nz.dropout(F.relu(_fc(W, "ps_net.%d" % idx, h)), site + j)
    out = _fc(W, "ps_net.9", h)
    return out[:, :S_DIM], out[:, S_DIM:]


def reparameterize(mean, logvar, nz, site):
    """ModelMid/ModelDown.reparameterize — src/torchmodel.py:54-56,130-132."""
    eps = nz.randn_like(mean, site)
    return eps * torch.exp(logvar * 0.5) + mean


def ps_forward_with_sample(W, pi, s0, nz, site):
    """ModelMid.transition_with_sample — src/torchmodel.py:63-66."""
    mean, logvar = ps_forward(W, pi, s0, nz, site)
    return reparameterize(mean, logvar, nz, site + 3), mean, logvar


def po_forward(W, s, nz, site):
    """ModelDown.decoder — src/torchmodel.py:106-128,139-141.  Dropout follows all four FCs."""
    h = s
    for j, idx in enumerate((0, 3, 6, 9)):
        h = nz.dropout(F.relu(_fc(W, "po_net.%d" % idx, h)), site + j)
    h = h.view(-1, 64, 16, 16)
    h = F.relu(F.conv_transpose2d(h, W["po_net.13.weight"], W["po_net.13.bias"], stride=1, padding=1))
    h = F.relu(F.conv_transpose2d(h, W["po_net.15.weight"], W["po_net.15.bias"], stride=2, padding=1,
                                  output_padding=1))
    h = F.relu(F.conv_transpose2d(h, W["po_net.17.weight"], W["po_net.17.bias"], stride=2, padding=1,
                                  output_padding=1))
    h = F.conv_transpose2d(h, W["po_net.19.weight"], W["po_net.19.bias"], stride=1, padding=1)
    return torch.sigmoid(h)


def qs_forward(W, o, nz, site):
    """ModelDown.encoder — src/torchmodel.py:84-104,134-137 with FC1 = 576->256 (D1)."""
    h = o
    for idx in (0, 2, 4, 6):
        h = F.relu(F.conv2d(h, W["qs_net.%d.weight" % idx], W["qs_net.%d.bias" % idx], stride=2))
    h = h.flatten(1)
    for j, idx in enumerate((9, 12, 15)):
        h = nz.dropout(F.relu(_fc(W, "qs_net.%d" % idx, h)), site + j)
    out = _fc(W, "qs_net.18", h)
    return out[:, :S_DIM], out[:, S_DIM:]


def qs_forward_with_sample(W, o, nz, site):
    """ModelDown.encoder_with_sample — src/torchmodel.py:143-146."""
    mean, logvar = qs_forward(W, o, nz, site)
    return reparameterize(mean, logvar, nz, site + 3), mean, logvar


# --------------------------------------------------------------------------- scalar terms

def entropy_normal_from_logvar(logvar):
    """src/torchutils.py:19-20."""
    return 0.5 * (LOG_2_PI_E + logvar)


def entropy_bernoulli(p):
    """src/torchutils.py:22-23 — note (delta + 1) - p is evaluated left to right in fp32."""
    return -(1 - p) * torch.log(DISPLACEMENT + 1 - p) - p * torch.log(DISPLACEMENT + p)


def check_reward(o):
    """ActiveInferenceModel.check_reward (64 px) — src/torchmodel.py:210-212 over
    src/torchutils.py:26-37, including the D12 broadcast: target[c,h,0] = 1[h < 32] for
    three identical channels against the whole NCHW image."""
    target = torch.zeros((3, 64, 1), dtype=torch.float32)
    target[:, :32] = 1.0
    x = o[:, 0:3, 0:64, :]
    lb = x * torch.log(DISPLACEMENT + target) + (1 - x) * torch.log(DISPLACEMENT + 1 - target)
    return torch.mean(lb, dim=[1, 2, 3]) * 10.0


# --------------------------------------------------------------------------- EFE evaluators

def calculate_G(W, s0, pi0, samples, nz, step=0, extras=None):
    """ActiveInferenceModel.calculate_G — src/torchmodel.py:270-300.
    Returns (G, [term0, term1, term2], ps1, ps1_mean, po1); the last three come from the
    LAST loop-2a sample.  `extras`, if a dict, receives term2_1 / term2_2 / ps1_logvar."""
    B = s0.shape[0]
    term0 = torch.zeros(B)
    term1 = torch.zeros(B)
    for j in range(samples):                                   # :273-281
        nz.at(step, j)
        ps1, ps1_mean, ps1_logvar = ps_forward_with_sample(W, pi0, s0, nz, SITES["PS_A"])
        po1 = po_forward(W, ps1, nz, SITES["PO_A"])
        _, _, qs1_logvar = qs_forward_with_sample(W, po1, nz, SITES["QS_A"])
        term0 += check_reward(po1)
        term1 += -torch.sum(entropy_normal_from_logvar(ps1_logvar)
                            + entropy_normal_from_logvar(qs1_logvar), dim=1)
    term0 /= float(samples)
    term1 /= float(samples)

    term2_1 = torch.zeros(B)
    term2_2 = torch.zeros(B)
    for j in range(samples):                                   # :287-292
        nz.at(step, j)
        s_fresh = ps_forward_with_sample(W, pi0, s0, nz, SITES["PS_B"])[0]
        term2_1 += torch.sum(entropy_bernoulli(po_forward(W, s_fresh, nz, SITES["PO_B1"])), dim=[1, 2, 3])
        s_rep = reparameterize(ps1_mean, ps1_logvar, nz, SITES["RP_B"])
        term2_2 += torch.sum(entropy_bernoulli(po_forward(W, s_rep, nz, SITES["PO_B2"])), dim=[1, 2, 3])
    term2_1 /= float(samples)
    term2_2 /= float(samples)
    term2 = term2_1 - term2_2
    G = -term0 + term1 + term2
    if extras is not None:
        extras.update(term2_1=term2_1, term2_2=term2_2, ps1_logvar=ps1_logvar)
    return G, [term0, term1, term2], ps1, ps1_mean, po1


def calculate_G_shard(W, s0, pi0, samples, j0, j1, nz, step=0):
    """The sample-sharded form of calculate_G (SURVEY.md §8 e): this "rank" evaluates samples [j0, j1) and
    returns the raw float64 sums (4,B) of term0, term1, term2_1, term2_2 over them — the all-reduce payload —
    plus (ps1, ps1_mean, ps1_logvar) of the GLOBALLY last loop-2a sample, which every rank recomputes under
    the same noise key (src/torchmodel.py:291,300 depend on it)."""
    B = s0.shape[0]
    sums = torch.zeros(4, B, dtype=torch.float64)
    nz.at(step, samples - 1)
    last = ps_forward_with_sample(W, pi0, s0, nz, SITES["PS_A"])
    for j in range(j0, j1):
        nz.at(step, j)
        ps1, ps1_mean, ps1_logvar = ps_forward_with_sample(W, pi0, s0, nz, SITES["PS_A"])
        po1 = po_forward(W, ps1, nz, SITES["PO_A"])
        _, _, qs1_logvar = qs_forward_with_sample(W, po1, nz, SITES["QS_A"])
        sums[0] += check_reward(po1).double()
        sums[1] += (-torch.sum(entropy_normal_from_logvar(ps1_logvar)
                               + entropy_normal_from_logvar(qs1_logvar), dim=1)).double()
    for j in range(j0, j1):
        nz.at(step, j)
        s_fresh = ps_forward_with_sample(W, pi0, s0, nz, SITES["PS_B"])[0]
        sums[2] += torch.sum(entropy_bernoulli(po_forward(W, s_fresh, nz, SITES["PO_B1"])), dim=[1, 2, 3]).double()
        s_rep = reparameterize(last[1], last[2], nz, SITES["RP_B"])
        sums[3] += torch.sum(entropy_bernoulli(po_forward(W, s_rep, nz, SITES["PO_B2"])), dim=[1, 2, 3]).double()
    return sums, last


def calculate_G_mean(W, s0, pi0, nz, step=0, extras=None):
    """ActiveInferenceModel.calculate_G_mean — src/torchmodel.py:302-327 (4-tuple).
    All 25 noises are still drawn; the eps of the two transitions is discarded."""
    nz.at(step, 0)
    _, ps1_mean, ps1_logvar = ps_forward_with_sample(W, pi0, s0, nz, SITES["PS_A"])
    po1 = po_forward(W, ps1_mean, nz, SITES["PO_A"])
    _, _, qs1_logvar = qs_forward_with_sample(W, po1, nz, SITES["QS_A"])
    term0 = check_reward(po1)
    term1 = -torch.sum(entropy_normal_from_logvar(ps1_logvar) + entropy_normal_from_logvar(qs1_logvar), dim=1)
    mean_b = ps_forward_with_sample(W, pi0, s0, nz, SITES["PS_B"])[1]
    term2_1 = torch.sum(entropy_bernoulli(po_forward(W, mean_b, nz, SITES["PO_B1"])), dim=[1, 2, 3])
    s_rep = reparameterize(ps1_mean, ps1_logvar, nz, SITES["RP_B"])
    term2_2 = torch.sum(entropy_bernoulli(po_forward(W, s_rep, nz, SITES["PO_B2"])), dim=[1, 2, 3])
    term2 = term2_1 - term2_2
    G = -term0 + term1 + term2
    if extras is not None:
        extras.update(term2_1=term2_1, term2_2=term2_2, ps1_logvar=ps1_logvar)
    return G, [term0, term1, term2], ps1_mean, po1


def calculate_G_repeated(W, o, pi, steps, calc_mean, samples, nz, four=False, trace=None):
    """calculate_G_repeated (src/torchmodel.py:227-245) and, with four=True,
    calculate_G_4_repeated (:247-268: pi = eye(4), calculate_G_mean when calc_mean)."""
    nz.at(0, 0)
    qs0_mean, qs0_logvar = qs_forward(W, o, nz, SITES["QS_ROOT"])
    qs0 = reparameterize(qs0_mean, qs0_logvar, nz, SITES["QS_ROOT"] + 3)
    B = o.shape[0]
    if four:
        pi = torch.eye(4)
    sum_terms = [torch.zeros(B) for _ in range(3)]
    sum_G = torch.zeros(B)
    s0 = qs0_mean if calc_mean else qs0
    po1 = None
    for t in range(steps):
        ex = {} if trace is not None else None
        if four and calc_mean:
            G, terms, ps1_mean, po1 = calculate_G_mean(W, s0, pi, nz, step=t, extras=ex)
            s1 = None
        else:
            G, terms, s1, ps1_mean, po1 = calculate_G(W, s0, pi, samples, nz, step=t, extras=ex)
        for i in range(3):
            sum_terms[i] += terms[i]
        sum_G += G
        if trace is not None:
            trace.append(dict(G=G.clone(), terms=[x.clone() for x in terms], s0=s0.clone(), **ex))
        s0 = ps1_mean if calc_mean else s1
    return sum_G, sum_terms, po1


def calculate_G_given_trajectory(W, s0_traj, ps1_traj, ps1_mean_traj, ps1_logvar_traj, pi0_traj, nz):
    """src/torchmodel.py:329-352."""
    nz.at(0, 0)
    po1 = po_forward(W, ps1_traj, nz, SITES["PO_A"])
    _, _, qs1_logvar = qs_forward_with_sample(W, po1, nz, SITES["QS_A"])
    term0 = check_reward(po1)
    term1 = -torch.sum(entropy_normal_from_logvar(ps1_logvar_traj) + entropy_normal_from_logvar(qs1_logvar), dim=1)
    s_fresh = ps_forward_with_sample(W, pi0_traj, s0_traj, nz, SITES["PS_B"])[0]
    term2_1 = torch.sum(entropy_bernoulli(po_forward(W, s_fresh, nz, SITES["PO_B1"])), dim=[1, 2, 3])
    s_rep = reparameterize(ps1_mean_traj, ps1_logvar_traj, nz, SITES["RP_B"])
    term2_2 = torch.sum(entropy_bernoulli(po_forward(W, s_rep, nz, SITES["PO_B2"])), dim=[1, 2, 3])
    return -term0 + term1 + (term2_1 - term2_2)


def _simulate_rollout(W, starting_s, depth, use_means, nz_roll):
    """The habit-policy rollout of src/torchmodel.py:354-388: `depth` B=1 transitions, noise cursor step = t."""
    s0 = torch.zeros((depth, S_DIM))
    ps1 = torch.zeros((depth, S_DIM))
    ps1_mean = torch.zeros((depth, S_DIM))
    ps1_logvar = torch.zeros((depth, S_DIM))
    pi0 = torch.zeros((depth, PI_DIM))
    s0[0] = starting_s
    qpi_ret = None
    for t in range(depth):
        nz_roll.at(t, 0)
        q = qpi_forward(W, s0[t].unsqueeze(0))[1][0]
        ok = bool(torch.isfinite(q).all() and (q >= 0).all() and q.sum() > 0)
        a = nz_roll.categorical(q, SITES["CAT"]) if ok else 0
        pi0[t, a] = 1.0
        if t == 0:
            qpi_ret = q if ok else pi0[0].clone()
        new, mean, logvar = ps_forward_with_sample(W, pi0[t].unsqueeze(0), s0[t].unsqueeze(0), nz_roll, SITES["PS_A"])
        ps1[t], ps1_mean[t], ps1_logvar[t] = new[0], mean[0], logvar[0]
        if t + 1 < depth:
            s0[t + 1] = mean[0] if use_means else new[0]
    return s0, ps1, ps1_mean, ps1_logvar, pi0, qpi_ret


def mcts_step_simulate(W, starting_s, depth, use_means, nz_roll, nz_traj):
    """src/torchmodel.py:354-393.  Habit-policy rollout of `depth` B=1 transitions (noise
    cursor: step = t, row 0), then calculate_G_given_trajectory over the depth rows under
    the next call key.  A categorical draw that cannot be made (NaN / negative / zero-sum
    probabilities, the reference's bare `except`, :365-367,380-381) falls back to action 0."""
    s0, ps1, ps1_mean, ps1_logvar, pi0, qpi_ret = _simulate_rollout(W, starting_s, depth, use_means, nz_roll)
    G = torch.mean(calculate_G_given_trajectory(W, s0, ps1, ps1_mean, ps1_logvar, pi0, nz_traj)).item()
    return G, pi0, qpi_ret


def mcts_step_simulate_batch(W, starts, depth, use_means, nz_roll, nz_traj):
    """K simulations in one pass (include/dai_b200.h dai_mcts_simulate_batch): rollout k is the
    reference's rollout drawn with noise row k; the K*depth trajectory rows (row = k*depth + t) go through ONE
    calculate_G_given_trajectory call; G[k] = mean over trajectory k.  K = 1 is mcts_step_simulate."""
    starts = starts.reshape(-1, S_DIM)
    parts = []
    for k in range(starts.shape[0]):
        nz_roll.row0 = k
        parts.append(_simulate_rollout(W, starts[k], depth, use_means, nz_roll))
    nz_roll.row0 = 0
    s0, ps1, mean, logvar, pi0 = (torch.cat([p[i] for p in parts]) for i in range(5))
    G = calculate_G_given_trajectory(W, s0, ps1, mean, logvar, pi0, nz_traj).reshape(-1, depth).mean(dim=1)
    return G, pi0.reshape(-1, depth, PI_DIM), torch.stack([p[5] for p in parts])


def select_actions(sum_G, temperature, nz):
    """The action choice of make_batch_dsprites_active_inference — src/util.py:46-53 (softmax_multi_with_log on
    -sum_G, including its un-tempered logSM) and :66-68, with the categorical draw taken from the keyed noise
    (row = root) instead of numpy's global generator."""
    x = (-sum_G.detach().numpy().astype(np.float32)).reshape(-1, 4)
    x = x - np.max(x, 1).reshape(-1, 1)
    e_x = np.exp(x / np.float32(temperature))
    SM = e_x / e_x.sum(axis=1).reshape(-1, 1)
    logSM = x - np.log(e_x.sum(axis=1).reshape(-1, 1) + np.float32(1e-20))
    nz.at(0, 0)
    choices = np.array([nz.categorical(torch.from_numpy(SM[r]), SITES["CAT"], row=r) for r in range(SM.shape[0])], dtype=np.int32)
    return SM, logSM, choices


# --------------------------------------------------------------------------- reference-shaped model

class _Sub:
    pass


class OracleModel:
    """CPU stand-in with the reference's ActiveInferenceModel attribute surface
    (src/torchmodel.py:149-165, used by src/mcts.py:71-85,158-164,188), backed by the
    functions above and keyed noise with the engine's call-counter convention:
    every API call uses key = seed + call_index, then call_index += 1
    (mcts_step_simulate uses two)."""

    def __init__(self, weights, seed=1234, call=0, training=True, noise="philox"):
        self.W = to_torch(weights) if not isinstance(next(iter(weights.values())), torch.Tensor) else weights
        self.s_dim, self.pi_dim = S_DIM, PI_DIM
        self.device = torch.device("cpu")
        self.precision = torch.float32
        self.pi_one_hot = torch.eye(4)
        self.pi_one_hot_3 = torch.eye(3)
        self.seed, self.call, self.training, self.kind = int(seed), int(call), training, noise
        self.tape = None
        md, mm, mt = _Sub(), _Sub(), _Sub()
        md.resolution = 64
        md.encoder = lambda o: qs_forward(self.W, o, self._nz(), SITES["QS_ROOT"])
        md.encoder_with_sample = lambda o: qs_forward_with_sample(self.W, o, self._nz(), SITES["QS_ROOT"])
        md.decoder = lambda s: po_forward(self.W, s, self._nz(), SITES["PO_A"])
        mm.transition = lambda pi, s0: ps_forward(self.W, pi, s0, self._nz(), SITES["PS_A"])
        mm.transition_with_sample = lambda pi, s0: ps_forward_with_sample(self.W, pi, s0, self._nz(), SITES["PS_A"])
        mt.encode_s = lambda s: qpi_forward(self.W, s)
        self.model_down, self.model_mid, self.model_top = md, mm, mt

    def _nz(self):
        if self.kind == "torch":
            return TorchStreamNoise(self.training)
        nz = PhiloxNoise(self.seed + self.call, self.training, self.tape)
        self.call += 1
        return nz

    def to(self, device):
        return self

    def set_rng(self, seed, call=0):
        self.seed, self.call = int(seed), int(call)

    def set_training(self, flag):
        self.training = bool(flag)

    def check_reward(self, o):
        return check_reward(o)

    def calculate_G(self, s0, pi0, samples=10):
        return calculate_G(self.W, s0, pi0, samples, self._nz())

    def calculate_G_mean(self, s0, pi0):
        return calculate_G_mean(self.W, s0, pi0, self._nz())

    def calculate_G_repeated(self, o, pi, steps=1, calc_mean=False, samples=10):
        return calculate_G_repeated(self.W, o, pi, steps, calc_mean, samples, self._nz())

    def calculate_G_4_repeated(self, o, steps=1, calc_mean=False, samples=10):
        return calculate_G_repeated(self.W, o, None, steps, calc_mean, samples, self._nz(), four=True)

    def calculate_G_given_trajectory(self, s0, ps1, ps1_mean, ps1_logvar, pi0):
        return calculate_G_given_trajectory(self.W, s0, ps1, ps1_mean, ps1_logvar, pi0, self._nz())

    def mcts_step_simulate(self, starting_s, depth, use_means=False):
        return mcts_step_simulate(self.W, starting_s, depth, use_means, self._nz(), self._nz())

    def mcts_step_simulate_batch(self, starting_s, depth, use_means=False):
        return mcts_step_simulate_batch(self.W, torch.as_tensor(starting_s), depth, use_means, self._nz(), self._nz())

    def select_actions(self, o_roots, steps=1, samples=10, calc_mean=False, temperature=10.0):
        o = torch.as_tensor(o_roots).reshape(-1, 1, 64, 64)
        R = o.shape[0]
        G, terms, _ = self.calculate_G_repeated(o.repeat_interleave(4, dim=0), torch.eye(4).repeat(R, 1), steps=steps,
                                                calc_mean=calc_mean, samples=samples)
        SM, logSM, choices = select_actions(G, temperature, self._nz())
        return torch.from_numpy(choices), torch.from_numpy(SM), torch.from_numpy(logSM), G, terms

    def habitual_net(self, o):
        return qpi_forward(self.W, self.model_down.encoder(o)[0])[1]

    def imagine_future_from_o(self, o0, pi):
        s0 = self.model_down.encoder_with_sample(o0)[0]
        ps1 = self.model_mid.transition_with_sample(pi, s0)[0]
        return self.model_down.decoder(ps1)

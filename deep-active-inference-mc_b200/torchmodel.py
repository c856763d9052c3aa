"""Drop-in for the reference's src/torchmodel.py on the EFE-rollout path.

Same class names, constructor, attributes, method names, argument order and return arity
as src/torchmodel.py:10-393 (zfountas/deep-active-inference-mc @ d7e76d8), so that the
reference's planner (src/mcts.py:71-85,158-164,188), batch maker (src/util.py:62) and demo
(test_demo.py:56-57,136-137,150) resolve against it unchanged.  The three sub-modules stay
real nn.Modules holding the canonical fp32 parameters (state_dict keys and checkpoint files
are the reference's); every forward on the path is served by the CUDA library through
engine.Engine — there is no eager-PyTorch or CPU fallback.

Differences a caller can observe, all forced by reference defects:
  D1  the encoder's first FC is 576->256 (the shipped 256->256 cannot run on 64x64 input);
  D2  `precision` exists;  D3 `.to()` exists and returns self;  D7 outputs are detached;
  G / terms / Qpi / pi0 / rewards come back as CPU tensors because src/mcts.py combines them
  in place with CPU tensors (:27-29,82,93); latents and images stay on the GPU.
Noise is keyed Philox (seed + call index), not torch's global generator: `set_rng`.
"""
import pickle

import torch
import torch.nn as nn

from .engine import Engine, DaiError  # noqa: F401


class _Served(nn.Module):
    """nn.Module whose forwards are served by the owning ActiveInferenceModel's engine."""

    def _eng(self):
        owner = self.__dict__.get("_owner")
        if owner is None:
            raise DaiError("%s is not attached to an ActiveInferenceModel" % type(self).__name__)
        owner._sync()
        return owner._engine

    def reparameterize(self, mean, logvar):
        # src/torchmodel.py:54-56,130-132 (host-visible helper; the hot path samples in-kernel)
        eps = torch.randn_like(mean)
        return eps * torch.exp(logvar * 0.5) + mean


class ModelTop(_Served):
    """Habit net Qpi — src/torchmodel.py:10-31."""

    def __init__(self, s_dim, pi_dim):
        super().__init__()
        self.s_dim, self.pi_dim = s_dim, pi_dim
        self.qpi_net = nn.Sequential(nn.Linear(s_dim, 128), nn.ReLU(), nn.Linear(128, 128), nn.ReLU(),
                                     nn.Linear(128, pi_dim))

    def encode_s(self, s0):
        logits, q, logq = self._eng().habit(s0)
        owner = self.__dict__["_owner"]
        return owner._host(logits), owner._host(q), owner._host(logq)


class ModelMid(_Served):
    """Transition net Ps — src/torchmodel.py:34-66."""

    def __init__(self, s_dim, pi_dim):
        super().__init__()
        self.s_dim, self.pi_dim = s_dim, pi_dim
        self.ps_net = nn.Sequential(nn.Linear(pi_dim + s_dim, 512), nn.ReLU(), nn.Dropout(0.5),
                                    nn.Linear(512, 512), nn.ReLU(), nn.Dropout(0.5),
                                    nn.Linear(512, 512), nn.ReLU(), nn.Dropout(0.5),
                                    nn.Linear(512, s_dim * 2))

    def transition(self, pi, s0):
        mean, logvar, _ = self._eng().transition(pi, s0, sample=False)
        return mean, logvar

    def transition_with_sample(self, pi, s0):
        mean, logvar, s = self._eng().transition(pi, s0, sample=True)
        return s, mean, logvar


class ModelDown(_Served):
    """Encoder Qs and decoder Po — src/torchmodel.py:69-146 (FC1 = 576->256, D1)."""

    def __init__(self, s_dim, pi_dim, colour_channels, resolution):
        super().__init__()
        if resolution != 64 or colour_channels != 1:
            raise ValueError("Unknown resolution")   # the 32-px branch is dead in the reference (D11)
        self.s_dim, self.pi_dim = s_dim, pi_dim
        self.colour_channels, self.resolution = colour_channels, resolution
        self.qs_net = nn.Sequential(
            nn.Conv2d(colour_channels, 32, kernel_size=3, stride=2), nn.ReLU(),
            nn.Conv2d(32, 32, kernel_size=3, stride=2), nn.ReLU(),
            nn.Conv2d(32, 64, kernel_size=3, stride=2), nn.ReLU(),
            nn.Conv2d(64, 64, kernel_size=3, stride=2), nn.ReLU(),
            nn.Flatten(),
            nn.Linear(64 * 3 * 3, 256), nn.ReLU(), nn.Dropout(0.5),
            nn.Linear(256, 256), nn.ReLU(), nn.Dropout(0.5),
            nn.Linear(256, 256), nn.ReLU(), nn.Dropout(0.5),
            nn.Linear(256, s_dim * 2))
        self.po_net = nn.Sequential(
            nn.Linear(s_dim, 256), nn.ReLU(), nn.Dropout(0.5),
            nn.Linear(256, 256), nn.ReLU(), nn.Dropout(0.5),
            nn.Linear(256, 256), nn.ReLU(), nn.Dropout(0.5),
            nn.Linear(256, 16 * 16 * 64), nn.ReLU(), nn.Dropout(0.5),
            nn.Unflatten(1, (64, 16, 16)),
            nn.ConvTranspose2d(64, 64, kernel_size=3, stride=1, padding=1), nn.ReLU(),
            nn.ConvTranspose2d(64, 64, kernel_size=3, stride=2, padding=1, output_padding=1), nn.ReLU(),
            nn.ConvTranspose2d(64, 32, kernel_size=3, stride=2, padding=1, output_padding=1), nn.ReLU(),
            nn.ConvTranspose2d(32, colour_channels, kernel_size=3, stride=1, padding=1), nn.Sigmoid())

    def encoder(self, o):
        mean, logvar, _ = self._eng().encode(o, sample=False)
        return mean, logvar

    def decoder(self, s):
        return self._eng().decode(s)

    def encoder_with_sample(self, o):
        mean, logvar, s = self._eng().encode(o, sample=True)
        return s, mean, logvar


class ActiveInferenceModel:
    """src/torchmodel.py:149-393."""

    def __init__(self, s_dim, pi_dim, gamma, beta_s, beta_o, colour_channels=1, resolution=64,
                 precision="bf16x3", device=None, seed=1234):
        if s_dim != 10 or pi_dim != 4:
            raise ValueError("the B200 path is built for s_dim=10, pi_dim=4 (BASELINE.json configs)")
        self._engine = Engine(device=device, precision=precision)
        self.device = self._engine.device
        self.precision = torch.float32                      # D2
        self.s_dim, self.pi_dim = s_dim, pi_dim
        self.model_top = ModelTop(s_dim, pi_dim).to(self.device)
        self.model_mid = ModelMid(s_dim, pi_dim).to(self.device)
        self.model_down = ModelDown(s_dim, pi_dim, colour_channels, resolution).to(self.device)
        for m in (self.model_top, self.model_mid, self.model_down):
            m.__dict__["_owner"] = self
        self.beta_s = torch.tensor(beta_s, device=self.device)
        self.gamma = torch.tensor(gamma, device=self.device)
        self.beta_o = torch.tensor(beta_o, device=self.device)
        self.pi_one_hot = torch.eye(4, device=self.device)
        self.pi_one_hot_3 = torch.eye(3, device=self.device)
        self.host_results = True
        self._versions = None
        self._plist = None
        self._train_flag = None
        self._group = None          # torch.distributed group for MC-sample sharding
        self._native_comm = False   # True: the C ABI owns the all-reduce (dai_comm_init)
        self._engine.set_rng(seed, 0)

    # ------------------------------------------------------------------ plumbing
    def to(self, device=None, *a, **k):     # D3: train.py:97 / test_demo.py:48 call .to(device)
        if device is not None and not isinstance(device, torch.dtype):
            d = torch.device(device)
            if d.type != "cuda" or (d.index is not None and d.index != self.device.index):
                raise DaiError("the model lives on %s (one engine handle per device); construct it with device=%s" % (self.device, d))
        return self

    def _entries(self):
        """[(state_dict key, owning module, parameter name)] — resolved once; the Parameter itself is looked up at
        every sync, so a parameter that is REPLACED (module.weight = nn.Parameter(...)) is seen like one that is
        updated in place."""
        if self._plist is None:
            self._plist = []
            for m in (self.model_top, self.model_mid, self.model_down):
                for name, _ in m.named_parameters():
                    mod_path, _, pname = name.rpartition(".")
                    self._plist.append((name, m.get_submodule(mod_path), pname))
        return self._plist

    def _sync(self):
        """Bring the engine's packed weights up to date: 46 (identity, version) reads when nothing changed (a few
        microseconds, no tensor work); otherwise ONLY the changed tensors are handed over (device-to-device) and only
        their packed images are rebuilt, on the device (dai_set_weight_async + dai_commit_weights)."""
        sig = self._versions
        dirty = None
        for i, (key, mod, pname) in enumerate(self._entries()):
            p = mod._parameters[pname]
            cur = (id(p), p._version, p.data_ptr())
            if sig is None or sig[i] != cur:
                if dirty is None:
                    dirty = {}
                    sig = list(sig) if sig is not None else [None] * len(self._plist)
                dirty[key] = p.detach()
                sig[i] = cur
        if dirty is not None:
            self._engine.set_weights(dirty)
            self._versions = sig
        # nn.Dropout follows module.training; the reference never leaves train mode.  One flag
        # serves the whole handle, so a call with the nets in different modes is refused instead of silently wrong.
        train = self.model_down.training
        if self.model_mid.training != train:
            raise DaiError("model_mid and model_down are in different train/eval modes; the engine applies one mode to all dropout sites")
        if train != self._train_flag:
            self._engine.set_training(train)
            self._train_flag = train

    def load_numpy_weights(self, weights):
        """Load {state_dict key: ndarray} (synthetic.make_weights) into the three modules."""
        for m in (self.model_top, self.model_mid, self.model_down):
            m.load_state_dict({k: torch.as_tensor(weights[k]) for k in m.state_dict()})
        self._versions = None
        return self

    def set_rng(self, seed, call=0):
        self._engine.set_rng(seed, call)

    def set_precision(self, precision):
        self._engine.set_precision(precision)

    def enable_sample_sharding(self, group=None, native=True):
        """Shard the MC samples of calculate_G / calculate_G_repeated over the ranks of a torch.distributed
        group: one all-reduce of the (4,B) float64 term sums per call.  native=True (default): the library owns
        the collective (dai_comm_init / dai_rollout_sharded: its own NCCL communicator, the all-reduce enqueued by
        the C ABI on the call's stream); torch.distributed only carries the 128-byte unique id from rank 0 to the
        others.  native=False: the sums come back and torch.distributed.all_reduce adds them."""
        import torch.distributed as dist
        self._group = group if group is not None else dist.group.WORLD
        self._native_comm = False
        if native:
            world, rank = dist.get_world_size(self._group), dist.get_rank(self._group)
            box = [self._engine.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=dist.get_global_rank(self._group, 0), group=self._group)
            self._engine.comm_init(box[0], rank, world)
            self._native_comm = True
        return self

    def _host(self, t):
        return t.cpu() if self.host_results else t

    def _shard(self, samples):
        if self._group is None:
            return None, 1
        import torch.distributed as dist
        from .sharding import shard_range
        world, rank = dist.get_world_size(self._group), dist.get_rank(self._group)
        return shard_range(samples, rank, world), world

    def _finish(self, out, samples, world):
        """all-reduce the raw term sums over the sample shards and finish G / terms."""
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(out["sums"], op=dist.ReduceOp.SUM, group=self._group)
            out["G"], out["t0"], out["t1"], out["t2"] = self._engine.combine(out["sums"], samples)
        return self._host(out["G"]), [self._host(out["t0"]), self._host(out["t1"]), self._host(out["t2"])]

    # ------------------------------------------------------------------ checkpoints (src/torchmodel.py:167-208)
    def save_weights(self, folder_chp):
        torch.save(self.model_down.state_dict(), f"{folder_chp}/checkpoint_down.pth")
        torch.save(self.model_top.state_dict(), f"{folder_chp}/checkpoint_top.pth")
        torch.save(self.model_mid.state_dict(), f"{folder_chp}/checkpoint_mid.pth")

    def load_weights(self, folder_chp):
        self.model_down.load_state_dict(torch.load(f"{folder_chp}/checkpoint_down.pth", map_location=self.device))
        self.model_top.load_state_dict(torch.load(f"{folder_chp}/checkpoint_top.pth", map_location=self.device))
        self.model_mid.load_state_dict(torch.load(f"{folder_chp}/checkpoint_mid.pth", map_location=self.device))
        self._versions = None

    def save_all(self, folder_chp, stats, script_file="", optimizers={}):
        self.save_weights(folder_chp)
        with open(f"{folder_chp}/stats.pkl", "wb") as ff:
            pickle.dump(stats, ff)
        with open(f"{folder_chp}/optimizers.pkl", "wb") as ff:
            pickle.dump({k: v.state_dict() for k, v in optimizers.items()}, ff)

    def load_all(self, folder_chp):
        self.load_weights(folder_chp)
        with open(f"{folder_chp}/stats.pkl", "rb") as ff:
            stats = pickle.load(ff)
        for name, key in (("beta_s", "var_beta_s"), ("gamma", "var_gamma"), ("beta_o", "var_beta_o")):
            if stats.get(key):
                setattr(self, name, torch.tensor(stats[key][-1], device=self.device))
        return stats, {}      # the reference always ends up with optimizers = {} (D4)

    # ------------------------------------------------------------------ small evaluators
    def check_reward(self, o):
        self._sync()
        return self._host(self._engine.check_reward(o))

    def imagine_future_from_o(self, o0, pi):
        s0, _, _ = self.model_down.encoder_with_sample(o0)
        ps1, _, _ = self.model_mid.transition_with_sample(pi, s0)
        return self.model_down.decoder(ps1)

    def habitual_net(self, o):
        qs_mean, _ = self.model_down.encoder(o)
        _, Qpi, _ = self.model_top.encode_s(qs_mean)
        return Qpi

    # ------------------------------------------------------------------ EFE evaluators
    def _hosted(self, out):
        return self._host(out["G"]), [self._host(out["t0"]), self._host(out["t1"]), self._host(out["t2"])]

    def calculate_G_repeated(self, o, pi, steps=1, calc_mean=False, samples=10):
        self._sync()
        if self._native_comm:
            out = self._engine.rollout_sharded(o, pi, steps, samples, calc_mean=calc_mean, four=False)
            return (*self._hosted(out), out["po1"])
        shard, world = self._shard(samples)
        out = self._engine.rollout(o, pi, steps, samples, calc_mean=calc_mean, four=False, shard=shard)
        G, terms = self._finish(out, samples, world)
        return G, terms, out["po1"]

    def select_actions(self, o_roots, steps=1, samples=10, calc_mean=False, temperature=10.0):
        """Batched many-roots action selection: the model-side core of
        make_batch_dsprites_active_inference (src/util.py:55-68) with the frames kept on the device.
        o_roots (R,1,64,64) or (R,64,64,1) -> (pi_choices (R,) int, Ppi (R,4), log_Ppi (R,4), sum_G (4R,), sum_terms)."""
        self._sync()
        o = self._engine.dev(o_roots).reshape(-1, 4096)
        R = o.shape[0]
        o4 = o.repeat_interleave(4, dim=0)                       # util.py:57  o0.repeat(4, 0)
        G, terms, _ = self.calculate_G_repeated(o4, torch.eye(4, device=self.device).repeat(R, 1), steps=steps,
                                                calc_mean=calc_mean, samples=samples)
        Ppi, logp, choice = self._engine.select_actions(G, temperature)
        return self._host(choice), self._host(Ppi), self._host(logp), G, terms

    def calculate_G_4_repeated(self, o, steps=1, calc_mean=False, samples=10):
        self._sync()
        if self._native_comm:
            out = self._engine.rollout_sharded(o, None, steps, samples, calc_mean=calc_mean, four=True)
            return (*self._hosted(out), out["po1"])
        if calc_mean:       # calculate_G_mean steps have a single sample: nothing to shard
            shard, world = None, 1
        else:
            shard, world = self._shard(samples)
        out = self._engine.rollout(o, None, steps, samples, calc_mean=calc_mean, four=True, shard=shard)
        G, terms = self._finish(out, samples, world)
        return G, terms, out["po1"]

    def calculate_G(self, s0, pi0, samples=10):
        self._sync()
        if self._native_comm:
            out = self._engine.calculate_G_sharded(s0, pi0, samples)
            return (*self._hosted(out), out["ps1"], out["ps1_mean"], out["po1"])
        shard, world = self._shard(samples)
        out = self._engine.calculate_G(s0, pi0, samples, shard=shard)
        G, terms = self._finish(out, samples, world)
        return G, terms, out["ps1"], out["ps1_mean"], out["po1"]

    def calculate_G_mean(self, s0, pi0):
        self._sync()
        out = self._engine.calculate_G_mean(s0, pi0)
        return (self._host(out["G"]), [self._host(out["t0"]), self._host(out["t1"]), self._host(out["t2"])],
                out["ps1_mean"], out["po1"])

    def calculate_G_given_trajectory(self, s0_traj, ps1_traj, ps1_mean_traj, ps1_logvar_traj, pi0_traj):
        self._sync()
        return self._host(self._engine.G_given_trajectory(s0_traj, ps1_traj, ps1_mean_traj, ps1_logvar_traj, pi0_traj))

    def mcts_step_simulate(self, starting_s, depth, use_means=False):
        self._sync()
        G, pi0, qpi = self._engine.mcts_simulate(starting_s, depth, use_means)
        return G, self._host(pi0), self._host(qpi)

    def mcts_step_simulate_batch(self, starting_s, depth, use_means=False):
        """K simulations in one pass: starting_s (K,10) -> (G (K,), pi0 (K,depth,4), Qpi (K,4)),
        all host tensors; K = 1 equals mcts_step_simulate."""
        self._sync()
        G, pi0, qpi = self._engine.mcts_simulate_batch(starting_s, depth, use_means)
        return G, self._host(pi0), self._host(qpi)

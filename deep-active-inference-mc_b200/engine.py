"""ctypes binding of the C ABI in include/dai_b200.h.

The library (libdai_b200.so, built in-tree by build.sh / __graft_entry__.build()) is the
only compute path: there is no CPU or eager-PyTorch fallback, and a missing library is a
hard error.  torch is used for device memory (tensor.data_ptr()), the current CUDA stream
and, in torchmodel.py, torch.distributed.
"""
import ctypes
import os

import torch

PREC_FP32_SIMT = 0
PREC_BF16X3 = 1
PREC_BF16X1 = 2
PRECISIONS = {"fp32_simt": PREC_FP32_SIMT, "bf16x3": PREC_BF16X3, "bf16x1": PREC_BF16X1}

_c_float_p = ctypes.c_void_p
_vp = ctypes.c_void_p


class DaiError(RuntimeError):
    pass


class DaiConfig(ctypes.Structure):
    _fields_ = [("s_dim", ctypes.c_int32), ("pi_dim", ctypes.c_int32), ("resolution", ctypes.c_int32),
                ("colour_channels", ctypes.c_int32), ("precision", ctypes.c_int32), ("training", ctypes.c_int32)]


class DaiMctsParams(ctypes.Structure):
    _fields_ = [("C", ctypes.c_float), ("threshold", ctypes.c_float), ("repeats", ctypes.c_int32),
                ("simulation_repeats", ctypes.c_int32), ("simulation_depth", ctypes.c_int32), ("use_means", ctypes.c_int32),
                ("using_prior_for_exploration", ctypes.c_int32), ("samples", ctypes.c_int32), ("leaves", ctypes.c_int32)]


class DaiMctsResult(ctypes.Structure):
    _fields_ = [("path_len", ctypes.c_int32), ("repeats_done", ctypes.c_int32), ("stopped", ctypes.c_int32),
                ("logged", ctypes.c_int32)]


class DaiStats(ctypes.Structure):
    _fields_ = [("kernel_launches", ctypes.c_uint64), ("calls", ctypes.c_uint64), ("workspace_bytes", ctypes.c_uint64),
                ("repack_launches", ctypes.c_uint64)]


# name -> (restype, argtypes); every symbol include/dai_b200.h declares
SIGNATURES = {
    "dai_create": (ctypes.c_int, [ctypes.POINTER(DaiConfig), ctypes.c_int, ctypes.POINTER(_vp)]),
    "dai_destroy": (ctypes.c_int, [_vp]),
    "dai_last_error": (ctypes.c_char_p, [_vp]),
    "dai_version": (ctypes.c_char_p, []),
    "dai_set_weight": (ctypes.c_int, [_vp, ctypes.c_char_p, _vp, ctypes.POINTER(ctypes.c_int64), ctypes.c_int]),
    "dai_set_weight_async": (ctypes.c_int, [_vp, ctypes.c_char_p, _vp, ctypes.POINTER(ctypes.c_int64), ctypes.c_int, _vp]),
    "dai_commit_weights": (ctypes.c_int, [_vp, _vp]),
    "dai_set_rng": (ctypes.c_int, [_vp, ctypes.c_uint64, ctypes.c_uint64]),
    "dai_get_rng": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]),
    "dai_set_training": (ctypes.c_int, [_vp, ctypes.c_int]),
    "dai_set_precision": (ctypes.c_int, [_vp, ctypes.c_int]),
    "dai_get_stats": (ctypes.c_int, [_vp, ctypes.POINTER(DaiStats), ctypes.c_int]),
    "dai_encode": (ctypes.c_int, [_vp, _vp, ctypes.c_int, _vp, _vp, _vp, _vp]),
    "dai_decode": (ctypes.c_int, [_vp, _vp, ctypes.c_int, _vp, _vp]),
    "dai_transition": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int, _vp, _vp, _vp, _vp]),
    "dai_habit": (ctypes.c_int, [_vp, _vp, ctypes.c_int, _vp, _vp, _vp, _vp]),
    "dai_check_reward": (ctypes.c_int, [_vp, _vp, ctypes.c_int, _vp, _vp]),
    "dai_calculate_G": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dai_calculate_G_mean": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dai_G_given_trajectory": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_int, _vp, _vp]),
    "dai_rollout": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dai_combine": (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp, _vp]),
    "dai_rollout_host": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dai_mcts_simulate": (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float),
                                         _vp, _vp, _vp]),
    "dai_mcts_simulate_batch": (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                               ctypes.POINTER(ctypes.c_float), _vp, _vp, _vp]),
    "dai_frames_set_sprites": (ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.POINTER(ctypes.c_int32), _vp]),
    "dai_frames_render": (ctypes.c_int, [_vp, _vp, ctypes.c_int, _vp, ctypes.c_int, ctypes.c_int, _vp,
                                         ctypes.POINTER(ctypes.c_int32), _vp]),
    "dai_mcts_plan": (ctypes.c_int, [_vp, _vp, _vp, ctypes.POINTER(DaiMctsParams), ctypes.POINTER(DaiMctsResult),
                                     ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32),
                                     ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_float), _vp]),
    "dai_select_actions": (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_float, _vp, _vp, _vp, _vp]),
    "dai_comm_unique_id": (ctypes.c_int, [_vp]),
    "dai_comm_init": (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_int]),
    "dai_comm_destroy": (ctypes.c_int, [_vp]),
    "dai_comm_info": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    "dai_rollout_sharded": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dai_calculate_G_sharded": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp,
                                               _vp, _vp, _vp, _vp, _vp]),
    "dai_profile_begin": (ctypes.c_int, [_vp]),
    "dai_profile_end": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int64),
                                       ctypes.POINTER(ctypes.c_int64), _vp]),
    "dai_debug_layer": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _vp, ctypes.c_int, _vp, _vp]),
}

_LIB = None


def library_path():
    """The in-tree build.  DAI_B200_LIB (A/B measurements of two builds only) names another build of the same sources."""
    return os.environ.get("DAI_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libdai_b200.so")


def load_library():
    """Load libdai_b200.so and bind every declared entry point (fails loudly if absent)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise DaiError("%s not found: build it first (./build.sh or __graft_entry__.build()); "
                       "there is no CPU fallback for this path" % path)
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype, fn.argtypes = res, args
    _LIB = lib
    return lib


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class Engine:
    """One handle on one CUDA device.  All tensor arguments are contiguous float32 CUDA tensors on
    that device unless a method says `host`."""

    def __init__(self, device=None, precision="bf16x3", training=True):
        if not torch.cuda.is_available():
            raise DaiError("no CUDA device: the EFE rollout path is CUDA-only (sm_100a), there is no CPU fallback")
        self.lib = load_library()
        idx = None if device is None else torch.device(device).index      # "cuda" without an index = the current device
        self.device = torch.device("cuda", torch.cuda.current_device() if idx is None else idx)
        cfg = DaiConfig(10, 4, 64, 1, PRECISIONS[precision] if isinstance(precision, str) else int(precision),
                        1 if training else 0)
        h = _vp()
        rc = self.lib.dai_create(ctypes.byref(cfg), self.device.index, ctypes.byref(h))
        if rc != 0:
            raise DaiError("dai_create failed with code %d" % rc)
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.dai_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _ck(self, rc):
        if rc != 0:
            raise DaiError("dai error %d: %s" % (rc, self.lib.dai_last_error(self.h).decode()))

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def dev(self, t, shape=None):
        """contiguous float32 tensor on the engine's device (host inputs are copied)."""
        t = torch.as_tensor(t)
        t = t.detach().to(device=self.device, dtype=torch.float32).contiguous()
        if shape is not None:
            t = t.reshape(shape)
        return t

    def new(self, *shape, dtype=torch.float32):
        return torch.empty(shape, device=self.device, dtype=dtype)

    # ------------------------------------------------------------------ state
    def set_weights(self, state):
        """state: {key: tensor / ndarray} — all 46 state_dict keys the first time, afterwards any subset (the tensors
        that changed).  Device tensors are copied device-to-device on the current stream, host tensors through the
        blocking entry point; the commit re-packs only the images of the tensors given (on the device, same stream)."""
        st = self._stream()
        with torch.cuda.device(self.device):
            for key, val in state.items():
                t = torch.as_tensor(val).detach()
                if t.dtype != torch.float32 or not t.is_contiguous():
                    t = t.to(torch.float32).contiguous()
                shape = (ctypes.c_int64 * t.dim())(*t.shape)
                if t.is_cuda:
                    self._ck(self.lib.dai_set_weight_async(self.h, key.encode(), _p(t), shape, t.dim(), st))
                else:
                    self._ck(self.lib.dai_set_weight(self.h, key.encode(), _p(t), shape, t.dim()))
            self._ck(self.lib.dai_commit_weights(self.h, st))

    def set_rng(self, seed, call=0):
        self._ck(self.lib.dai_set_rng(self.h, int(seed), int(call)))

    def get_rng(self):
        s, c = ctypes.c_uint64(), ctypes.c_uint64()
        self._ck(self.lib.dai_get_rng(self.h, ctypes.byref(s), ctypes.byref(c)))
        return s.value, c.value

    def set_training(self, flag):
        self._ck(self.lib.dai_set_training(self.h, 1 if flag else 0))

    def set_precision(self, precision):
        self._ck(self.lib.dai_set_precision(self.h, PRECISIONS[precision] if isinstance(precision, str) else int(precision)))

    def stats(self, reset=False):
        s = DaiStats()
        self._ck(self.lib.dai_get_stats(self.h, ctypes.byref(s), 1 if reset else 0))
        return {"kernel_launches": s.kernel_launches, "calls": s.calls, "workspace_bytes": s.workspace_bytes,
                "repack_launches": s.repack_launches}

    # ------------------------------------------------------------------ nets
    def encode(self, o, sample=False):
        o = self.dev(o).reshape(-1, 4096)
        B = o.shape[0]
        mean, logvar = self.new(B, 10), self.new(B, 10)
        s = self.new(B, 10) if sample else None
        self._ck(self.lib.dai_encode(self.h, _p(o), B, _p(mean), _p(logvar), _p(s), self._stream()))
        return mean, logvar, s

    def decode(self, s):
        s = self.dev(s).reshape(-1, 10)
        B = s.shape[0]
        po = self.new(B, 1, 64, 64)
        self._ck(self.lib.dai_decode(self.h, _p(s), B, _p(po), self._stream()))
        return po

    def transition(self, pi, s0, sample=False):
        pi, s0 = self.dev(pi).reshape(-1, 4), self.dev(s0).reshape(-1, 10)
        B = s0.shape[0]
        if pi.shape[0] != B:
            raise DaiError("transition: pi has %d rows, s0 has %d" % (pi.shape[0], B))
        mean, logvar = self.new(B, 10), self.new(B, 10)
        s = self.new(B, 10) if sample else None
        self._ck(self.lib.dai_transition(self.h, _p(pi), _p(s0), B, _p(mean), _p(logvar), _p(s), self._stream()))
        return mean, logvar, s

    def habit(self, s):
        s = self.dev(s).reshape(-1, 10)
        B = s.shape[0]
        logits, q, logq = self.new(B, 4), self.new(B, 4), self.new(B, 4)
        self._ck(self.lib.dai_habit(self.h, _p(s), B, _p(logits), _p(q), _p(logq), self._stream()))
        return logits, q, logq

    def check_reward(self, o):
        o = self.dev(o).reshape(-1, 4096)
        r = self.new(o.shape[0])
        self._ck(self.lib.dai_check_reward(self.h, _p(o), o.shape[0], _p(r), self._stream()))
        return r

    # ------------------------------------------------------------------ EFE
    def calculate_G(self, s0, pi0, samples, shard=None, want_po1=True):
        """Returns dict(sums (4,B) f64, G, t0, t1, t2, ps1, ps1_mean, ps1_logvar, po1)."""
        s0, pi0 = self.dev(s0).reshape(-1, 10), self.dev(pi0).reshape(-1, 4)
        B = s0.shape[0]
        j0, j1 = (0, samples) if shard is None else shard
        out = dict(sums=self.new(4, B, dtype=torch.float64), G=self.new(B), t0=self.new(B), t1=self.new(B),
                   t2=self.new(B), ps1=self.new(B, 10), ps1_mean=self.new(B, 10), ps1_logvar=self.new(B, 10),
                   po1=self.new(B, 1, 64, 64) if want_po1 else None)
        self._ck(self.lib.dai_calculate_G(self.h, _p(s0), _p(pi0), B, samples, j0, j1, _p(out["sums"]), _p(out["G"]),
                                          _p(out["t0"]), _p(out["t1"]), _p(out["t2"]), _p(out["ps1"]),
                                          _p(out["ps1_mean"]), _p(out["ps1_logvar"]), _p(out["po1"]), self._stream()))
        return out

    def calculate_G_mean(self, s0, pi0):
        s0, pi0 = self.dev(s0).reshape(-1, 10), self.dev(pi0).reshape(-1, 4)
        B = s0.shape[0]
        out = dict(G=self.new(B), t0=self.new(B), t1=self.new(B), t2=self.new(B), ps1_mean=self.new(B, 10),
                   po1=self.new(B, 1, 64, 64))
        self._ck(self.lib.dai_calculate_G_mean(self.h, _p(s0), _p(pi0), B, _p(out["G"]), _p(out["t0"]), _p(out["t1"]),
                                               _p(out["t2"]), _p(out["ps1_mean"]), _p(out["po1"]), self._stream()))
        return out

    def G_given_trajectory(self, s0, ps1, ps1_mean, ps1_logvar, pi0):
        s0, ps1, mu, lv = (self.dev(x).reshape(-1, 10) for x in (s0, ps1, ps1_mean, ps1_logvar))
        pi0 = self.dev(pi0).reshape(-1, 4)
        D = s0.shape[0]
        G = self.new(D)
        self._ck(self.lib.dai_G_given_trajectory(self.h, _p(s0), _p(ps1), _p(mu), _p(lv), _p(pi0), D, _p(G), self._stream()))
        return G

    def rollout(self, o, pi, steps, samples, calc_mean=False, four=False, shard=None, want_po1=True):
        o = self.dev(o).reshape(-1, 4096)
        B = o.shape[0]
        pi = None if pi is None else self.dev(pi).reshape(B, 4)
        j0, j1 = (0, samples) if shard is None else shard
        out = dict(sums=self.new(4, B, dtype=torch.float64), G=self.new(B), t0=self.new(B), t1=self.new(B),
                   t2=self.new(B), po1=self.new(B, 1, 64, 64) if want_po1 else None)
        self._ck(self.lib.dai_rollout(self.h, _p(o), _p(pi), B, steps, samples, 1 if calc_mean else 0, 1 if four else 0,
                                      j0, j1, _p(out["sums"]), _p(out["G"]), _p(out["t0"]), _p(out["t1"]),
                                      _p(out["t2"]), _p(out["po1"]), self._stream()))
        return out

    # ------------------------------------------------------------------ sample sharding behind the C ABI
    def comm_unique_id(self):
        """128 bytes (ncclUniqueId) from rank 0, to be handed to every rank's comm_init."""
        buf = ctypes.create_string_buffer(128)
        rc = self.lib.dai_comm_unique_id(buf)
        if rc != 0:
            raise DaiError("dai_comm_unique_id failed with code %d (NCCL not loadable?)" % rc)
        return buf.raw

    def comm_init(self, unique_id, rank, world):
        with torch.cuda.device(self.device):
            self._ck(self.lib.dai_comm_init(self.h, ctypes.c_char_p(unique_id) if unique_id is not None else None, int(rank), int(world)))

    def comm_info(self):
        r, w = ctypes.c_int(), ctypes.c_int()
        self._ck(self.lib.dai_comm_info(self.h, ctypes.byref(r), ctypes.byref(w)))
        return r.value, w.value

    def rollout_sharded(self, o, pi, steps, samples, calc_mean=False, four=False, want_po1=True):
        """dai_rollout_sharded: this rank's sample slice + the library's own all-reduce + finish."""
        o = self.dev(o).reshape(-1, 4096)
        B = o.shape[0]
        pi = None if pi is None else self.dev(pi).reshape(B, 4)
        out = dict(G=self.new(B), t0=self.new(B), t1=self.new(B), t2=self.new(B),
                   po1=self.new(B, 1, 64, 64) if want_po1 else None)
        self._ck(self.lib.dai_rollout_sharded(self.h, _p(o), _p(pi), B, steps, samples, 1 if calc_mean else 0, 1 if four else 0,
                                              _p(out["G"]), _p(out["t0"]), _p(out["t1"]), _p(out["t2"]), _p(out["po1"]),
                                              self._stream()))
        return out

    def calculate_G_sharded(self, s0, pi0, samples, want_po1=True):
        s0, pi0 = self.dev(s0).reshape(-1, 10), self.dev(pi0).reshape(-1, 4)
        B = s0.shape[0]
        out = dict(G=self.new(B), t0=self.new(B), t1=self.new(B), t2=self.new(B), ps1=self.new(B, 10),
                   ps1_mean=self.new(B, 10), ps1_logvar=self.new(B, 10), po1=self.new(B, 1, 64, 64) if want_po1 else None)
        self._ck(self.lib.dai_calculate_G_sharded(self.h, _p(s0), _p(pi0), B, samples, _p(out["G"]), _p(out["t0"]), _p(out["t1"]),
                                                  _p(out["t2"]), _p(out["ps1"]), _p(out["ps1_mean"]), _p(out["ps1_logvar"]),
                                                  _p(out["po1"]), self._stream()))
        return out

    def combine(self, sums, samples):
        B = sums.shape[1]
        G, t0, t1, t2 = self.new(B), self.new(B), self.new(B), self.new(B)
        self._ck(self.lib.dai_combine(self.h, _p(sums), B, samples, _p(G), _p(t0), _p(t1), _p(t2), self._stream()))
        return G, t0, t1, t2

    def rollout_host(self, o_host, pi_host, steps, samples, calc_mean, four, out_host, po1_host=None):
        """o_host (B,4096) and out_host (4,B) are HOST float32 tensors (pinned for async copies);
        copies in, runs, copies G,t0,t1,t2 back and waits (the bench's end-to-end call)."""
        B = o_host.shape[0]
        self._ck(self.lib.dai_rollout_host(self.h, _p(o_host), _p(pi_host), B, steps, samples, 1 if calc_mean else 0,
                                           1 if four else 0, _p(out_host[0]), _p(out_host[1]), _p(out_host[2]),
                                           _p(out_host[3]), _p(po1_host), self._stream()))

    def select_actions(self, G, temperature=10.0):
        """G (4R,) -> Ppi (R,4), logPpi (R,4), choice (R,) int32 (src/util.py:46-53,66-68)."""
        G = self.dev(G).reshape(-1)
        R = G.shape[0] // 4
        Ppi, logp = self.new(R, 4), self.new(R, 4)
        choice = self.new(R, dtype=torch.int32)
        self._ck(self.lib.dai_select_actions(self.h, _p(G), R, float(temperature), _p(Ppi), _p(logp), _p(choice), self._stream()))
        return Ppi, logp, choice

    def profile_begin(self):
        self._ck(self.lib.dai_profile_begin(self.h))

    def profile_end(self):
        """{layer: (ms, launches, rows)} for layers fc4, ct1, ct2, ct3, pixel."""
        ms, n, rows = (ctypes.c_float * 5)(), (ctypes.c_int64 * 5)(), (ctypes.c_int64 * 5)()
        self._ck(self.lib.dai_profile_end(self.h, ms, n, rows, self._stream()))
        return {name: (ms[i], n[i], rows[i]) for i, name in enumerate(("fc4", "ct1", "ct2", "ct3", "pixel"))}

    def debug_layer(self, layer, precision, x):
        """test hook: one decoder contraction layer, fp32 NHWC in/out (include/dai_b200.h)."""
        x = self.dev(x)
        n = x.shape[0]
        prec = PRECISIONS[precision] if isinstance(precision, str) else int(precision)
        hw_out, c_out = {1: (256, 64), 2: (1024, 64), 3: (4096, 32), 23: (4096, 32)}[layer]
        # layer 3 on tensor cores (and 23, the fused ct2 -> ct3 pair kernel) returns the last deconv's row planes
        # [n][3][4096] (see dai_tc.cu), not the 32 channels
        out = self.new(n, 3, 4096) if (layer in (3, 23) and prec != PREC_FP32_SIMT) else self.new(n, hw_out, c_out)
        self._ck(self.lib.dai_debug_layer(self.h, layer, prec, _p(x), n, _p(out), self._stream()))
        return out

    def mcts_simulate(self, starting_s, depth, use_means=False):
        s = self.dev(starting_s).reshape(10)
        pi0, qpi = self.new(depth, 4), self.new(4)
        G = ctypes.c_float()
        self._ck(self.lib.dai_mcts_simulate(self.h, _p(s), depth, 1 if use_means else 0, ctypes.byref(G), _p(pi0),
                                            _p(qpi), self._stream()))
        return float(G.value), pi0, qpi

    def mcts_simulate_batch(self, starting_s, depth, use_means=False):
        """K habit-policy rollouts + one trajectory evaluation (include/dai_b200.h dai_mcts_simulate_batch).
        starting_s (K,10) -> G (K,) host float32 tensor, pi0 (K,depth,4), qpi (K,4) on the device."""
        s = self.dev(starting_s).reshape(-1, 10)
        K = s.shape[0]
        pi0, qpi = self.new(K, depth, 4), self.new(K, 4)
        G = (ctypes.c_float * K)()
        self._ck(self.lib.dai_mcts_simulate_batch(self.h, _p(s), K, depth, 1 if use_means else 0, G, _p(pi0), _p(qpi),
                                                  self._stream()))
        return torch.tensor(list(G), dtype=torch.float32), pi0, qpi

    # ------------------------------------------------------------------ frame producer
    def set_sprites(self, imgs, latents_sizes):
        """imgs: HOST uint8 (count,64,64[,1]) binary sprite table (dSprites `imgs`); latents_sizes: 6 ints."""
        imgs = torch.as_tensor(imgs).to(torch.uint8).reshape(-1, 4096).contiguous()
        if imgs.is_cuda:
            imgs = imgs.cpu()
        sizes = (ctypes.c_int32 * 6)(*[int(x) for x in latents_sizes])
        self._ck(self.lib.dai_frames_set_sprites(self.h, _p(imgs), imgs.shape[0], sizes, self._stream()))

    def render_frames(self, current_s, last_r, reference_bases=False, check=True):
        """Game.current_frame_all: current_s (G, >=6), last_r (G) -> (G,1,64,64) float32 on the device."""
        s = self.dev(current_s)
        r = self.dev(last_r).reshape(-1)
        G = s.shape[0]
        o = self.new(G, 1, 64, 64)
        bad = ctypes.c_int32(0)
        self._ck(self.lib.dai_frames_render(self.h, _p(s), s.shape[1], _p(r), G, 1 if reference_bases else 0, _p(o),
                                            ctypes.byref(bad) if check else None, self._stream()))
        if check and bad.value:
            raise ValueError("Error: %d game(s) with a sprite index outside the table or a reward outside [-1, 1]" % bad.value)
        return o

    def mcts_plan(self, frame, params, leaves, qs0_mean=None):
        """One planning decision with the tree on the device (include/dai_b200.h dai_mcts_plan).  Returns
        (raw_path, repeats_done, stopped, all_paths, all_paths_G)."""
        f = None if frame is None else self.dev(frame).reshape(4096)
        q = None if qs0_mean is None else self.dev(qs0_mean).reshape(10)
        prm = DaiMctsParams(float(params.C), float(params.threshold), int(params.repeats), int(params.simulation_repeats),
                            int(params.simulation_depth), 1 if params.use_means else 0,
                            1 if params.using_prior_for_exploration else 0, int(getattr(params, "samples", 1)), int(leaves))
        res = DaiMctsResult()
        n = max(int(params.repeats), 1)
        path = (ctypes.c_int32 * 64)()
        all_paths, all_len, all_G = (ctypes.c_int32 * (n * 64))(), (ctypes.c_int32 * n)(), (ctypes.c_float * n)()
        self._ck(self.lib.dai_mcts_plan(self.h, _p(f), _p(q), ctypes.byref(prm), ctypes.byref(res), path, all_paths, all_len,
                                        all_G, self._stream()))
        paths = [[int(all_paths[i * 64 + d]) for d in range(all_len[i])] for i in range(res.logged)]
        return ([int(path[i]) for i in range(res.path_len)], int(res.repeats_done), bool(res.stopped), paths,
                [float(all_G[i]) for i in range(res.logged)])

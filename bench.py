"""EFE-rollout throughput on B200 (BASELINE.json metric) — see the contract in DESIGN.md §Measurement.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one batch: R root observations x 4 actions, N MC
samples, horizon T (calculate_G_repeated layout, row = root*4 + action) = R rollouts.
Workload = BASELINE.json configs[1]: dSprites-like 64x64 frames, 4 actions, N=50, T=10.
`value` times the device-resident call (dai_rollout); `e2e` times dai_rollout_host with the
frames in pinned host memory and G/terms read back, every step.  Synthetic frames and
random-init weights (no dataset / checkpoint exists offline).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MAC_NODE_EVAL = 134_721_312          # 2 Ps + 3 Po + 1 Qs per (row, sample, step)  
MAC_QS = 3_868_960
MAC_CT2 = 9_437_184                  # ConvT 64->64, 16x16 -> 32x32, per decoder row
MAC_CT3 = 18_874_368                 # ConvT 64->32, 32x32 -> 64x64, per decoder row


def rollout_flops(n, t):
    return 2.0 * (4 * n * t * MAC_NODE_EVAL + 4 * MAC_QS)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(samples, horizon, steps, warmup, extra_as_shipped=1):
    """The reference's own CPU path timed on this box's host cores, at the stated config: every timed step is ONE
    full `calculate_G_4_repeated(o, steps=T, samples=N)` call for one root observation (test_demo.py:150) — a
    bounded sample (1 root) of the GPU arm's R-root step; roots are independent, so rollouts/s = 1 / seconds per call.
    Runs the UNMODIFIED reference (staged under baseline/_ref, or /root/reference where mounted; SHIM-1/2 of
    SURVEY.md §0.1 on the instance) with torch's own RNG; falls back to the oracle port over the same ATen kernels
    and draw pattern when the reference is absent.  `value` is measured under torch.no_grad() (the faster, i.e.
    conservative, figure); `extra_as_shipped` more calls are timed with autograd on, as the reference ships."""
    os.environ.setdefault("CUDA_VISIBLE_DEVICES", "")
    import torch
    from dai_b200 import synthetic
    from oracle import efe_oracle as O
    from oracle import reference_model as RM
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    w = synthetic.make_weights(0)
    o = torch.from_numpy(synthetic.make_frames(1, 0)).repeat(4, 1, 1, 1)
    if RM.available():
        ref = RM.load(w)
        kind, where = "reference", RM.location()

        def call():
            return ref.calculate_G_4_repeated(o, steps=horizon, samples=samples)
    else:
        W = O.to_torch(w)
        kind, where = "port", "oracle/efe_oracle.py"

        def call():
            return O.calculate_G_repeated(W, o, None, horizon, False, samples, O.TorchStreamNoise(), four=True)
    torch.manual_seed(0)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t = time.perf_counter()
            call()
            if i >= warmup:
                times.append(time.perf_counter() - t)
    shipped = []
    for i in range(extra_as_shipped):
        t = time.perf_counter()
        call()
        shipped.append(time.perf_counter() - t)
    per_call = sum(times) / len(times)
    return {"value": 1.0 / per_call, "per_call_s": per_call, "cores": torch.get_num_threads(), "kind": kind, "where": where,
            "as_shipped_value": (len(shipped) / sum(shipped)) if shipped else None}


def mcts_decisions(model, N, T, leaves, device_tree, n_timed, n_warm=1):
    """configs[3]: full planner, 30 expansions x N samples x simulation depth T; seconds per decision (host wall clock:
    the planner's host syncs are part of a decision)."""
    import torch
    from dai_b200 import synthetic
    from dai_b200 import mcts as planner
    prm = planner.MCTS_Params()
    prm.repeats, prm.threshold, prm.use_means, prm.samples, prm.simulation_depth = 30, 2.0, False, N, T
    frame = torch.from_numpy(synthetic.make_frames(1, 0))[0, 0]
    times = []
    for i in range(n_warm + n_timed):
        torch.cuda.synchronize()
        t = time.perf_counter()
        if device_tree:
            planner.active_inference_mcts_device(model, frame, prm, o_shape=(1, 64, 64), leaves=leaves)
        elif leaves > 1:
            planner.active_inference_mcts_batched(model, frame, prm, o_shape=(1, 64, 64), leaves=leaves)
        else:
            planner.active_inference_mcts(model, frame, prm, o_shape=(1, 64, 64))
        torch.cuda.synchronize()
        if i >= n_warm:
            times.append(time.perf_counter() - t)
    return sum(times) / len(times)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--roots", type=int, default=16, help="root observations per step per job (R)")
    ap.add_argument("--samples", type=int, default=50)
    ap.add_argument("--horizon", type=int, default=10)
    ap.add_argument("--precision", default="bf16x3")
    ap.add_argument("--shard", default="hybrid", choices=["hybrid", "samples", "roots"],
                    help="N>1: hybrid = pairs of ranks split the MC samples of 2R roots (one NCCL all-reduce per "
                         "rollout inside each pair), R*N roots in total [weak]; samples = all ranks split the "
                         "samples of the same R roots [strong]; roots = R independent roots per rank, no collective [weak]")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the configs[2..4] legs (`extra` in the JSON line)")
    ap.add_argument("--workload", default="rollout", choices=["rollout", "mcts"],
                    help="rollout = BASELINE.json configs[1] (default, the bench line); mcts = configs[3]: full planner, "
                         "30 expansions x N=50 samples x simulation depth 10, reported as decisions/s (extra line, N=1 only)")
    ap.add_argument("--leaves", type=int, default=1, help="--workload mcts: leaves expanded per batch (1 = the "
                    "reference's sequential search; >1 = batched-leaf planner, SURVEY.md §8 f2)")
    ap.add_argument("--device-tree", action="store_true", help="--workload mcts: search tree resident on the GPU "
                    "(dai_mcts_plan), one host wait per decision")
    ap.add_argument("--quick", action="store_true", help="profiling pass: 1 warm-up, no e2e / cpu / extra legs (never a bench value)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    N, T, R = args.samples, args.horizon, args.roots
    workload = "dSprites-like 64x64, 4 actions, N_samples=%d, horizon T=%d (BASELINE.json configs[1])" % (N, T)
    config = {"workload": workload, "roots_per_step": R, "rows_per_step": 4 * R, "samples": N, "horizon": T}

    if args.impl == "reference":
        if rank != 0:
            return
        k, w = max(1, args.steps), max(0, args.warmup)
        r = cpu_reference_run(N, T, k, w)
        sample = ("each step = one full calculate_G_4_repeated(o, steps=%d, samples=%d) call for 1 of the %d roots of the "
                  "GPU arm's step (roots are independent); %d timed + %d warm-up calls under torch.no_grad(), %.2f s each; "
                  "as shipped (autograd on): %.4f rollouts/s; code: %s" %
                  (T, N, R, k, w, r["per_call_s"], r["as_shipped_value"] or 0.0, r["where"]))
        print(json.dumps({
            "impl": "reference", "metric": "EFE rollouts/sec", "value": r["value"], "unit": "rollouts/s",
            "n_gpus": args.gpus, "steps": k, "warmup": w, "ms_per_step": r["per_call_s"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config,
            "cpu_baseline": {"value": r["value"], "unit": "rollouts/s", "cores": r["cores"], "kind": r["kind"], "sample": sample,
                             "as_shipped_value": r["as_shipped_value"]},
            "e2e": {"value": r["value"], "unit": "rollouts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    import torch
    import torch.distributed as dist
    from dai_b200 import synthetic
    from dai_b200.sharding import shard_range
    from dai_b200.torchmodel import ActiveInferenceModel

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    model = ActiveInferenceModel(10, 4, 1.0, 1.0, 1.0, precision=args.precision, device=dev)
    model.load_numpy_weights(synthetic.make_weights(0))
    model._sync()
    eng = model._engine

    if args.workload == "mcts":
        eng.stats(reset=True)
        dt = mcts_decisions(model, N, T, args.leaves, args.device_tree, args.steps, max(1, args.warmup))
        print(json.dumps({"metric": "MCTS decisions/sec", "value": 1.0 / dt, "unit": "decisions/s", "n_gpus": 1,
                          "steps": args.steps, "warmup": max(1, args.warmup), "ms_per_step": dt * 1e3,
                          "higher_is_better": True, "data": "synthetic", "dtype": args.precision,
                          "expansions_per_s": 31.0 / dt,
                          "config": {"workload": "full MCTS, 30 expansions x N=%d samples, simulation depth %d "
                                                 "(BASELINE.json configs[3])" % (N, T), "leaves_per_batch": args.leaves, "tree": "device" if args.device_tree else "host",
                                     "timing": "host wall clock, "
                                     "2 host syncs per expansion are inherent in the planner (src/mcts.py:82,188)"},
                          "gpu_launches": int(eng.stats()["kernel_launches"])}))
        return

    # work split over ranks
    group, frame_seed = None, 0
    if world > 1 and args.shard == "roots":
        my_roots, shard, scaling = R, None, "weak"          # every rank owns R roots: R*world rollouts per step
        total_roots = R * world
        frame_seed = rank
    elif world > 1 and args.shard == "hybrid":
        ws = 2 if world % 2 == 0 else 1                      # sample-shard degree (50 samples -> 25 + 25)
        groups = [dist.new_group(ranks=[g * ws + i for i in range(ws)]) for g in range(world // ws)]
        group, frame_seed = groups[rank // ws], rank // ws
        my_roots, shard, scaling = R * ws, (shard_range(N, rank % ws, ws) if ws > 1 else None), "weak"
        total_roots = R * world                              # per-GPU work is constant: R roots x N samples x T
    else:
        my_roots, shard, scaling = R, (shard_range(N, rank, world) if world > 1 else None), ("strong" if world > 1 else "weak")
        total_roots = R
    flush = torch.empty(160 * 1024 * 1024 // 4, device=dev)      # > 126 MB L2

    def make_steps(roots, n, t, shard_, group_, seed, native=False):
        """(device-resident step, end-to-end step, h2d bytes, d2h bytes) for `roots` roots x 4 actions, n samples, horizon t.
        Both return what the reference call returns: G, the three terms and po1 of the last step.  native: the sample
        shards are combined inside the C ABI (dai_rollout_sharded: the library's own NCCL all-reduce)."""
        frames = torch.from_numpy(synthetic.make_frames(roots, seed=seed))
        o_host = frames.repeat_interleave(4, dim=0).reshape(4 * roots, 4096).contiguous().pin_memory()
        o_dev = o_host.to(dev)
        out_host = torch.empty(4, 4 * roots).pin_memory()
        po1_host = torch.empty(4 * roots, 4096).pin_memory()

        def step_device():
            if native:
                return eng.rollout_sharded(o_dev, None, t, n, calc_mean=False, four=False, want_po1=True)
            out = eng.rollout(o_dev, None, t, n, calc_mean=False, four=False, shard=shard_, want_po1=True)
            if shard_ is not None:
                dist.all_reduce(out["sums"], group=group_)
                eng.combine(out["sums"], n)
            return out

        def step_e2e():
            if shard_ is None and not native:
                eng.rollout_host(o_host, None, t, n, False, False, out_host, po1_host)
            else:
                o = o_host.to(dev, non_blocking=True)
                if native:
                    out = eng.rollout_sharded(o, None, t, n, calc_mean=False, four=False, want_po1=True)
                    G, t0, t1, t2 = out["G"], out["t0"], out["t1"], out["t2"]
                else:
                    out = eng.rollout(o, None, t, n, calc_mean=False, four=False, shard=shard_, want_po1=True)
                    dist.all_reduce(out["sums"], group=group_)
                    G, t0, t1, t2 = eng.combine(out["sums"], n)
                out_host[0].copy_(G, non_blocking=True)
                out_host[1].copy_(t0, non_blocking=True)
                out_host[2].copy_(t1, non_blocking=True)
                out_host[3].copy_(t2, non_blocking=True)
                po1_host.copy_(out["po1"].reshape(4 * roots, 4096), non_blocking=True)
                torch.cuda.current_stream().synchronize()

        return step_device, step_e2e, int(o_host.numel() * 4), int((out_host.numel() + po1_host.numel()) * 4)

    def timed(fn, k, w):
        for _ in range(w):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
        for a, b in ev:
            flush.zero_()                                   # L2 flush between timed iterations (not timed)
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = sum(a.elapsed_time(b) for a, b in ev)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / k

    step_device, step_e2e, h2d, d2h = make_steps(my_roots, N, T, shard, group, frame_seed)
    sampler = ClockSampler(local)
    eng.stats(reset=True)
    if rank == 0:
        sampler.start()
    ms_dev = timed(step_device, args.steps, 1 if args.quick else max(3, args.warmup))
    launches = eng.stats()["kernel_launches"]
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = ms_dev if args.quick else timed(step_e2e, args.steps, 1)
    value = total_roots / (ms_dev * 1e-3)
    e2e = total_roots / (ms_e2e * 1e-3)
    # roofline of the dominant kernel (ct3: ConvT 64->32 as tcgen05 implicit GEMM): every launch of the same
    # steps bracketed by CUDA events on the launch stream, in a separate pass so the events do not perturb
    # `value` (all ranks run it: the sharded step contains the all-reduce)
    layers = {}
    if args.precision != "fp32_simt":
        eng.profile_begin()
        for _ in range(min(args.steps, 3)):
            step_device()
        layers = eng.profile_end()

    # ---- the other BASELINE.json configs, timed in the same run (`extra`) -------------------------------------
    extra = {}
    if not args.quick and not args.no_extras:
        ks = max(2, min(args.steps, 5))
        if world == 1:
            # single-root latency of configs[1] (test_demo.py:150 is R=1); first of the extra legs, 10 timed calls, after a 2 s
            # pause: a lone request meets an idle GPU, not one still at the clock the power cap left after the main leg
            time.sleep(2.0)
            d1, e1, _, _ = make_steps(1, N, T, None, None, 2)
            ms1 = timed(d1, max(ks, 10), 5)
            ms1e = timed(e1, max(ks, 10), 1)
            extra["r1_latency_ms"] = {"value": ms1, "e2e": ms1e, "unit": "ms per rollout", "rollouts_s": 1e3 / ms1,
                                      "config": "configs[1] with R=1: one root (4 action rows), N=%d, T=%d; measured after a 2 s pause" % (N, T)}
            # configs[2]: N=200, T=10 (4 roots per step)
            d3, e3, _, _ = make_steps(4, 200, 10, None, None, 1)
            ms3 = timed(d3, ks, 2)
            ms3e = timed(e3, ks, 1)
            extra["c3_rollouts_s"] = {"value": 4 / (ms3 * 1e-3), "e2e": 4 / (ms3e * 1e-3), "unit": "rollouts/s", "ms_per_step": ms3,
                                      "config": "configs[2]: N=200, T=10, R=4 roots per step",
                                      "algorithmic_tflops": rollout_flops(200, 10) * 4 / (ms3 * 1e-3) / 1e12}
            # configs[3]: full MCTS decision (host wall clock)
            dseq = mcts_decisions(model, 50, 10, 1, False, 3)
            dtree = mcts_decisions(model, 50, 10, 16, True, 5)
            extra["c4_decisions_s"] = {"sequential": 1.0 / dseq, "device_tree_leaves16": 1.0 / dtree, "unit": "decisions/s",
                                       "ms_sequential": dseq * 1e3, "ms_device_tree_leaves16": dtree * 1e3,
                                       "config": "configs[3]: 30 expansions x N=50 samples x simulation depth 10; sequential = the "
                                                 "reference's search order, one leaf per expansion (src/mcts.py:150-195); "
                                                 "device_tree = dai_mcts_plan, 16 leaves per batch"}
        # configs[4]: N=800 samples sharded over ALL ranks, T=15, one root, ONE all-reduce of the (4,B) f64 term sums over
        # the world group per rollout (at N=1: the same rollout unsharded, the strong-scaling base)
        sh5 = shard_range(800, rank, world) if world > 1 else None
        if world > 1:
            model.enable_sample_sharding(native=True)        # dai_comm_init over all ranks: the C ABI owns the all-reduce
        d5, e5, _, _ = make_steps(1, 800, 15, sh5, None, 3, native=world > 1)
        ms5 = timed(d5, ks, 2)
        ms5e = timed(e5, ks, 1)
        # the same partition with 16 roots per step (rows per rank as in the main line): how the 8-rank all-reduce partition
        # scales when every rank has a full batch
        d5b, _, _, _ = make_steps(16, 800, 15, sh5, None, 4, native=world > 1)
        ms5b = timed(d5b, 2, 1)
        extra["c5_sample_sharded_R16"] = {"value": 16.0 / (ms5b * 1e-3), "unit": "rollouts/s", "ms_per_step": ms5b, "n_gpus": world,
                                          "scaling": "strong", "config": "configs[4] with 16 roots per step: N=800 samples over all ranks, T=15",
                                          "algorithmic_tflops": 16 * rollout_flops(800, 15) / (ms5b * 1e-3) / 1e12}
        extra["c5_sample_sharded"] = {"value": 1.0 / (ms5 * 1e-3), "e2e": 1.0 / (ms5e * 1e-3), "unit": "rollouts/s", "ms_per_step": ms5,
                                      "n_gpus": world, "scaling": "strong", "samples_per_rank": (sh5[1] - sh5[0]) if sh5 else 800,
                                      "collective": ("one ncclAllReduce(sum) of (4,4) float64 over %d ranks per rollout, issued inside "
                                                     "dai_rollout_sharded (C ABI)" % world) if world > 1 else "none",
                                      "config": "configs[4]: N=800 samples, T=15, R=1 root",
                                      "algorithmic_tflops": rollout_flops(800, 15) / (ms5 * 1e-3) / 1e12}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf, peak_src = (peaks.get("bf16_tflops_sustained"), "measured (bf16_tflops_sustained)") if peaks else (1400.0, "fallback")
    flops_step = rollout_flops(N, T) * total_roots
    line = {
        "metric": "EFE rollouts/sec", "value": value, "unit": "rollouts/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_dev, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": {"fp32_simt": "f32", "bf16x3": "bf16x3 (hi/lo split, f32 accumulate)",
                                       "bf16x1": "bf16"}[args.precision],
        "data": "synthetic", "config": dict(config, precision=args.precision, shard=args.shard if world > 1 else "none",
                                            roots_total=total_roots, roots_per_rank=my_roots,
                                            samples_per_rank=(shard[1] - shard[0]) if shard else N,
                                            returns="G, term0..2 and po1 of the last step, like the reference call",
                                            l2="flushed between timed iterations (160 MB write)"),
        "node_evals_per_s": value * 4 * N * T, "algorithmic_tflops": flops_step / (ms_dev * 1e-3) / 1e12,
        "e2e": {"value": e2e, "unit": "rollouts/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    # dominant kernel: the fused ct2 -> ct3 pair kernel (k_tc_ct23; both layers are timed as "ct2" and "ct3" has no launches),
    # or ct3 alone when the layers run separately (DAI_TC_FUSE23=0)
    fused = bool(layers) and layers["ct3"][1] == 0 and layers["ct2"][1] > 0
    dom = "ct2" if fused else "ct3"
    mac_dom = (MAC_CT2 + MAC_CT3) if fused else MAC_CT3
    roof = {"bound": "tensor", "achieved": None, "peak": peak_tf, "unit": "TFLOP/s", "frac": None, "traffic": None,
            "peak_source": peak_src,
            "kernel": "k_tc_ct23 (ConvT 64->64 16x16->32x32 + ConvT 64->32 32x32->64x64 fused on CTA pairs, + last-deconv projection)"
                      if fused else "k_tc_conv<ct3> (ConvT 64->32, 32x32->64x64, + last-deconv projection)"}
    if layers and layers[dom][1] > 0:
        ms3, n3, rows3 = layers[dom]
        alg = 2.0 * mac_dom * rows3 / (ms3 * 1e-3) / 1e12           # algorithmic flops: MAC per decoder row
        issued = alg * (3 if args.precision == "bf16x3" else 1)     # bf16x3 issues 3 MMAs per algorithmic MAC
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ct23_dram_traffic.json" if fused else "ct3_dram_traffic.json")))["bytes_per_launch"]
        except Exception:
            pass
        roof.update({"achieved": alg, "frac": alg / peak_tf, "issued_mma_tflops": issued, "issued_frac": issued / peak_tf,
                     "avg_launch_ms": ms3 / n3, "launches_timed": int(n3), "rows_per_launch": rows3 / n3,
                     "traffic": traffic, "mac_per_row": mac_dom,
                     "step_share_ms": {k: v[0] / min(args.steps, 3) for k, v in layers.items()}})
        # the other contraction kernels of the decoder, same accounting (algorithmic MAC per decoder row)
        macs = {"fc4": 4_194_304, "ct1": 9_437_184, "ct2": mac_dom if fused else MAC_CT2, "ct3": MAC_CT3}
        nprod = 3 if args.precision == "bf16x3" else 1
        roof["kernels"] = {("ct2+ct3" if (fused and k == "ct2") else k):
                           {"alg_tflops": 2.0 * m * layers[k][2] / (layers[k][0] * 1e-3) / 1e12,
                            "issued_frac": nprod * 2.0 * m * layers[k][2] / (layers[k][0] * 1e-3) / 1e12 / peak_tf}
                           for k, m in macs.items() if k in layers and layers[k][0] > 0}
        roof["whole_step"] = {"alg_tflops": flops_step / (ms_dev * 1e-3) / 1e12 / world,
                              "issued_frac": nprod * flops_step / (ms_dev * 1e-3) / 1e12 / world / peak_tf}
    line["roofline"] = roof
    if extra:
        line["extra"] = extra
    if not args.no_cpu_baseline and not args.quick and world == 1:
        try:
            r = cpu_reference_run(N, T, 3, 1)        # 4 full calls + 1 as shipped: ~15-20 s of CPU work on the box
            line["cpu_baseline"] = {"value": r["value"], "unit": "rollouts/s", "cores": r["cores"], "kind": r["kind"],
                                    "as_shipped_value": r["as_shipped_value"],
                                    "sample": "3 timed + 1 warm-up full calculate_G_4_repeated(o, steps=%d, samples=%d) calls for 1 root "
                                              "under torch.no_grad() (%.2f s each), +1 as shipped; code: %s" % (T, N, r["per_call_s"], r["where"])}
        except Exception as e:                       # never lose the measured line to the baseline leg
            line["cpu_baseline"] = {"value": None, "unit": "rollouts/s", "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (e,)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

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
MAC_CT3 = 18_874_368                 # dominant kernel: ConvT 64->32, 32x32 -> 64x64, per decoder row


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


def cpu_reference_run(samples, horizon, steps, warmup):
    """The reference's CPU path (oracle port over the same ATen kernels, torch's own RNG draws as in the
    reference — SURVEY.md §6) timed on this box's host cores.  One step = one bounded sample:
    calculate_G_4_repeated(o, steps=1, samples=N) for one root; CPU cost is linear in N*T
    (BASELINE.md §2), so rollouts/s at horizon T = 1 / (T * seconds per 1-step call)."""
    os.environ.setdefault("CUDA_VISIBLE_DEVICES", "")
    import torch
    from dai_b200 import synthetic
    from oracle import efe_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    W = O.to_torch(synthetic.make_weights(0))
    o = torch.from_numpy(synthetic.make_frames(1, 0)).repeat(4, 1, 1, 1)
    torch.manual_seed(0)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t = time.perf_counter()
            O.calculate_G_repeated(W, o, None, 1, False, samples, O.TorchStreamNoise(), four=True)
            if i >= warmup:
                times.append(time.perf_counter() - t)
    per_call = sum(times) / len(times)
    return 1.0 / (per_call * horizon), per_call, torch.get_num_threads()


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
    ap.add_argument("--workload", default="rollout", choices=["rollout", "mcts"],
                    help="rollout = BASELINE.json configs[1] (default, the bench line); mcts = configs[3]: full planner, "
                         "30 expansions x N=50 samples x simulation depth 10, reported as decisions/s (extra line, N=1 only)")
    ap.add_argument("--leaves", type=int, default=1, help="--workload mcts: leaves expanded per batch (1 = the "
                    "reference's sequential search; >1 = batched-leaf planner, SURVEY.md §8 f2)")
    ap.add_argument("--device-tree", action="store_true", help="--workload mcts: search tree resident on the GPU "
                    "(dai_mcts_plan), one host wait per decision")
    ap.add_argument("--quick", action="store_true", help="profiling pass: 1 warm-up, no e2e / cpu legs (never a bench value)")
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
        k = max(1, min(args.steps, 20))
        value, per_call, cores = cpu_reference_run(N, T, k, max(1, min(args.warmup, 2)))
        sample = "calculate_G_4_repeated(steps=1, samples=%d), 1 root, %d timed calls, scaled by 1/T (cost is linear in N*T)" % (N, k)
        print(json.dumps({
            "impl": "reference", "metric": "EFE rollouts/sec", "value": value, "unit": "rollouts/s",
            "n_gpus": args.gpus, "steps": k, "warmup": max(1, min(args.warmup, 2)), "ms_per_step": per_call * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(config, roots_per_step=1, rows_per_step=4),
            "cpu_baseline": {"value": value, "unit": "rollouts/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "rollouts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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
        from dai_b200 import mcts as planner
        prm = planner.MCTS_Params()
        prm.repeats, prm.threshold, prm.use_means, prm.samples, prm.simulation_depth = 30, 2.0, False, N, T
        frame = torch.from_numpy(synthetic.make_frames(1, 0))[0, 0]
        eng.stats(reset=True)
        times = []
        for i in range(max(1, args.warmup) + args.steps):
            torch.cuda.synchronize()
            t = time.perf_counter()
            if args.device_tree:
                planner.active_inference_mcts_device(model, frame, prm, o_shape=(1, 64, 64), leaves=args.leaves)
            elif args.leaves > 1:
                planner.active_inference_mcts_batched(model, frame, prm, o_shape=(1, 64, 64), leaves=args.leaves)
            else:
                planner.active_inference_mcts(model, frame, prm, o_shape=(1, 64, 64))
            torch.cuda.synchronize()
            if i >= max(1, args.warmup):
                times.append(time.perf_counter() - t)
        dt = sum(times) / len(times)
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
    frames = torch.from_numpy(synthetic.make_frames(my_roots, seed=frame_seed))
    o_host = frames.repeat_interleave(4, dim=0).reshape(4 * my_roots, 4096).contiguous().pin_memory()
    o_dev = o_host.to(dev)
    out_host = torch.empty(4, 4 * my_roots).pin_memory()
    flush = torch.empty(160 * 1024 * 1024 // 4, device=dev)      # > 126 MB L2

    def step_device():
        out = eng.rollout(o_dev, None, T, N, calc_mean=False, four=False, shard=shard, want_po1=False)
        if shard is not None:
            dist.all_reduce(out["sums"], group=group)
            eng.combine(out["sums"], N)
        return out

    def step_e2e():
        if shard is None:
            eng.rollout_host(o_host, None, T, N, False, False, out_host)
        else:
            o = o_host.to(dev, non_blocking=True)
            out = eng.rollout(o, None, T, N, calc_mean=False, four=False, shard=shard, want_po1=False)
            dist.all_reduce(out["sums"], group=group)
            G, t0, t1, t2 = eng.combine(out["sums"], N)
            out_host[0].copy_(G, non_blocking=True)
            out_host[1].copy_(t0, non_blocking=True)
            out_host[2].copy_(t1, non_blocking=True)
            out_host[3].copy_(t2, non_blocking=True)
            torch.cuda.current_stream().synchronize()

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
                                            l2="flushed between timed iterations (160 MB write)"),
        "node_evals_per_s": value * 4 * N * T, "algorithmic_tflops": flops_step / (ms_dev * 1e-3) / 1e12,
        "e2e": {"value": e2e, "unit": "rollouts/s", "h2d_bytes_per_step": int(o_host.numel() * 4),
                "d2h_bytes_per_step": int(out_host.numel() * 4)},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    roof = {"bound": "tensor", "achieved": None, "peak": peak_tf, "unit": "TFLOP/s", "frac": None, "traffic": None,
            "peak_source": peak_src, "kernel": "k_tc_conv<ct3> (ConvT 64->32, 32x32->64x64, + last-deconv projection)"}
    if layers and layers["ct3"][1] > 0:
        ms3, n3, rows3 = layers["ct3"]
        alg = 2.0 * MAC_CT3 * rows3 / (ms3 * 1e-3) / 1e12           # algorithmic flops of ct3: 18,874,368 MAC per decoder row
        issued = alg * (3 if args.precision == "bf16x3" else 1)     # bf16x3 issues 3 MMAs per algorithmic MAC
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ct3_dram_traffic.json")))["bytes_per_launch"]
        except Exception:
            pass
        roof.update({"achieved": alg, "frac": alg / peak_tf, "issued_mma_tflops": issued, "issued_frac": issued / peak_tf,
                     "avg_launch_ms": ms3 / n3, "launches_timed": int(n3), "rows_per_launch": rows3 / n3,
                     "traffic": traffic,
                     "step_share_ms": {k: v[0] / min(args.steps, 3) for k, v in layers.items()}})
        # the other contraction kernels of the decoder, same accounting (algorithmic MAC per decoder row)
        macs = {"fc4": 4_194_304, "ct1": 9_437_184, "ct2": 9_437_184, "ct3": MAC_CT3}
        nprod = 3 if args.precision == "bf16x3" else 1
        roof["kernels"] = {k: {"alg_tflops": 2.0 * m * layers[k][2] / (layers[k][0] * 1e-3) / 1e12,
                               "issued_frac": nprod * 2.0 * m * layers[k][2] / (layers[k][0] * 1e-3) / 1e12 / peak_tf}
                           for k, m in macs.items() if k in layers and layers[k][0] > 0}
    line["roofline"] = roof
    if not args.no_cpu_baseline and not args.quick and world == 1:
        ncalls = 24                                   # ~10 s of CPU work on the box's host cores (0.35 s per call)
        cv, per_call, cores = cpu_reference_run(N, T, ncalls, 2)
        line["cpu_baseline"] = {"value": cv, "unit": "rollouts/s", "cores": cores, "kind": "port",
                                "sample": "calculate_G_4_repeated(steps=1, samples=%d), 1 root, %d timed calls "
                                          "(%.2f s each), scaled by 1/T" % (N, ncalls, per_call)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Benchmark of the MSACL fused rollout hot path (BASELINE.json metric: fused env-steps/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--env NAME] [--envs-per-gpu M]

Workload (config 5 of BASELINE.json, per-GPU share): QuadTracking, 2^21 env instances per GPU
(16 Mi envs on 8 GPUs), default-init StochaPolicy actor (seed 0), Philox resets, one "step" =
one fused rollout launch of `--inner` (16) env steps over all instances, transitions
materialised in HBM.  Weak scaling: per-GPU work is fixed, env ids are sharded by rank.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "fused rollout env-steps/sec"
UNIT = "env-steps/s"


def actor_flops(d, a):      # 2*(D*256 + 256*256 + 256*2A)
    return 2 * (d * 256 + 256 * 256 + 256 * 2 * a)


DYN_FLOPS = {"VanderPol": 70, "Pendulum": 70, "DuctedFan": 200, "TwoLink": 270, "SingleTrackCar": 430, "QuadTracking": 950}


def transition_bytes(d, a):  # obs, act, rew, cost, obs2, logp (f32) + done, emit (u8) per env-step
    return 4 * (2 * d + a + 3) + 2


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 4), ("hw_thermal_slowdown", 5), ("sw_thermal_slowdown", 6), ("sw_power_cap", 7)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the same fused step (actor + sample + env + reward/cost +
# autoreset) on the host cores -- bench.py's cpu_baseline leg and --impl reference
# ------------------------------------------------------------------------------------------
def cpu_rollout(env_name, n_envs, steps, seed=0):
    from oracle import actor as oactor, envs as oenv, rollout as oroll
    spec = oenv.SPECS[env_name]
    w = oactor.init_policy_weights(spec.obs_dim, spec.act_dim, seed=0)
    ids = np.arange(n_envs, dtype=np.uint64)
    venv = oroll.VectorEnv(env_name, oroll.philox_reset(env_name, seed, ids, np.zeros(n_envs, np.int64)), seed=seed, env_ids=ids)
    rng = np.random.default_rng(seed)
    t0 = time.perf_counter()
    for _ in range(steps):
        eps = rng.standard_normal((n_envs, spec.act_dim)).astype(np.float32)
        oroll.sampler_step(venv, w, eps)
    return time.perf_counter() - t0


def _cpu_worker(job):
    env_name, n_envs, steps, seed = job
    try:
        from threadpoolctl import threadpool_limits
        with threadpool_limits(limits=1):                 # one BLAS thread per process: the processes cover the cores
            return cpu_rollout(env_name, n_envs, steps, seed)
    except ImportError:
        return cpu_rollout(env_name, n_envs, steps, seed)


def cpu_rollout_all_cores(env_name, n_per_proc, steps, procs):
    """P independent processes (no IPC on the step path), each stepping its own n_per_proc env instances with the
    NumPy port -- the embarrassingly-parallel best case of BASELINE.md section 3.  Returns env-steps/s over the
    slowest process."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        times = pool.map(_cpu_worker, [(env_name, n_per_proc, steps, 1000 + p) for p in range(procs)])
    return procs * n_per_proc * steps / max(times)


def cpu_baseline(env_name, budget_s=12.0):
    n, procs = 4096, os.cpu_count() or 1
    t1 = cpu_rollout(env_name, n, 1)                      # includes first-call overheads
    t2 = cpu_rollout(env_name, n, 2)
    per = max((t2 - t1), 1e-3) * 1.5                      # processes slow each other down a little
    steps = int(min(max(budget_s / per, 2), 200))
    val = cpu_rollout_all_cores(env_name, n, steps, procs)
    return {"value": val, "unit": UNIT, "cores": procs, "kind": "port",
            "sample": f"{env_name}: {procs} processes x {n} envs x {steps} fused steps (NumPy oracle port, one process per host core)"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, steps_per, procs = 4096, 4, os.cpu_count() or 1
    for _ in range(args.warmup):
        cpu_rollout(args.env, n, 1)
    t0 = time.perf_counter()
    vals = [cpu_rollout_all_cores(args.env, n, steps_per, procs) for _ in range(args.steps)]
    el = time.perf_counter() - t0
    val = float(np.mean(vals))
    sample = f"{args.env}: each step = {procs} processes x {n} envs x {steps_per} fused env steps on the NumPy oracle port"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.env} fused rollout (actor+sample+env+reward/cost+autoreset), bounded CPU sample of config 5"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--env", default="QuadTracking")
    ap.add_argument("--envs-per-gpu", type=int, default=1 << 21)
    ap.add_argument("--inner", type=int, default=16, help="env steps per fused launch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--replay-batch", type=int, default=256)
    ap.add_argument("--e2e-full-d2h", action="store_true",
                    help="also time a variant that copies EVERY transition of the chunk to pinned host memory")
    ap.add_argument("--engine", default="tc", choices=["tc", "ffma"],
                    help="actor engine: tcgen05 split-bf16 tensor cores (default) or FP32 FFMA")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist

    import msacl_b200  # noqa: F401
    from msacl_b200 import _lib
    from msacl_b200.buffer import B200NstepReplayBuffer
    from msacl_b200.sampler import ActorWeights, FusedRollout
    from msacl_b200.specs import get_spec
    from msacl_b200 import distributed as mdist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    spec = get_spec(args.env)
    n, K, n_step = args.envs_per_gpu, args.inner, 20
    D, A = spec.obs_dim, spec.act_dim

    # actor: torch default nn.Linear init, seed 0 (random-init weights of the reference architecture)
    torch.manual_seed(0)
    pol = torch.nn.Sequential(torch.nn.Linear(D, 256), torch.nn.ReLU(), torch.nn.Linear(256, 256), torch.nn.ReLU(),
                              torch.nn.Linear(256, 2 * A))
    host_w = [p.detach().clone().pin_memory() for p in pol.parameters()]
    h2d_bytes = sum(p.numel() * 4 for p in host_w)

    def upload_actor():
        dw = [p.to(dev, non_blocking=True) for p in host_w]
        return ActorWeights([(dw[0], dw[1]), (dw[2], dw[3]), (dw[4], dw[5])], device=dev)

    ro = FusedRollout(args.env, n, K, n_step=n_step, seed=0, env_base=rank * n, device=dev, engine=args.engine)
    ro.state.reset()
    actor = upload_actor()
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (`value`) + dominant-kernel timing
    for _ in range(args.warmup):
        ro.run(actor)
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record(stream)
    for i in range(args.steps):
        ev[i][0].record(stream)
        ro.run(actor)
        ev[i][1].record(stream)
    t_end.record(stream)
    barrier()
    clk = clocks.stop()
    total_ms = t_start.elapsed_time(t_end)
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    if world > 1:
        tt = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    steps_total = n * K * args.steps * world
    value = steps_total / (total_ms * 1e-3)

    # ---------------- end to end through the public API with host buffers
    buf = B200NstepReplayBuffer(obs_dim=D, act_dim=A, buffer_max_size=1_000_000, n_step=n_step, device=dev)
    Bq = args.replay_batch // world if world > 1 else args.replay_batch
    host_batch = {k: torch.empty(Bq * world, n_step, *v.shape[2:], dtype=torch.float32).pin_memory() for k, v in buf.n_step_buf.items()}
    host_stats = torch.empty(8, dtype=torch.float64).pin_memory()
    d2h_bytes = sum(v.numel() * 4 for v in host_batch.values()) + 64

    def e2e_step():
        a = upload_actor()                               # H2D: the learner's current policy (pinned host memory)
        batch = ro.run(a)                                # fused K-step rollout
        buf.add_batch(batch)                             # n-step windows -> device replay ring
        sub = buf.sample_batch(Bq)                       # replay batch for the learner
        stats = ro.stats[:8].clone()
        if world > 1:                                    # NCCL: replay-batch all-gather + episode statistics sum, one bucket
            sub, stats = mdist.exchange_batch_and_stats(sub, stats)
        for k in host_batch:                             # D2H: the replay batch + statistics
            host_batch[k].copy_(sub[k], non_blocking=True)
        host_stats.copy_(stats, non_blocking=True)
        stream.synchronize()

    for _ in range(args.warmup):
        e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        e2e_step()
    e1.record(stream)
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item())
    e2e_value = steps_total / (e2e_ms * 1e-3)

    # ---------------- optional: the same step but with every transition of the chunk copied to the host
    e2e_full = None
    if args.e2e_full_d2h:
        tr = ro.tr
        dev_fields = {k: v[tr.H:] for k, v in tr.fields().items()}
        host_fields = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in dev_fields.items()}
        full_bytes = sum(v.numel() * v.element_size() for v in host_fields.values())

        def full_step():
            a = upload_actor()
            ro.run(a)
            cur = tr.fields()                            # views of the chunk this launch wrote
            for k in host_fields:
                host_fields[k].copy_(cur[k][tr.H:], non_blocking=True)
            stream.synchronize()

        for _ in range(2):
            full_step()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for _ in range(args.steps):
            full_step()
        f1.record(stream)
        barrier()
        full_ms = f0.elapsed_time(f1)
        e2e_full = {"value": steps_total / (full_ms * 1e-3), "unit": UNIT, "ms_per_step": full_ms / args.steps,
                    "d2h_bytes_per_step": full_bytes, "what": "rollout + D2H of all K x n transition records (pinned host memory)"}

    # ---------------- roofline of the dominant kernel (rollout_fused_kernel): FP32 FFMA pipe
    flops_per_step = actor_flops(D, A) + DYN_FLOPS[args.env]
    achieved_tflops = flops_per_step * n * K / (kern_ms * 1e-3) / 1e12
    sink = torch.rand(128, device=dev)
    import ctypes as C
    lib = _lib.load()

    def probe(mode, iters):
        fl = C.c_double(0.0)
        for _ in range(2):
            _lib.check(lib.msacl_ffma_probe(mode, iters, sink.data_ptr(), C.byref(fl), _lib.current_stream()))
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(stream)
        _lib.check(lib.msacl_ffma_probe(mode, iters, sink.data_ptr(), C.byref(fl), _lib.current_stream()))
        p1.record(stream)
        torch.cuda.synchronize()
        return fl.value / (p0.elapsed_time(p1) * 1e-3) / 1e12

    ffma_peak = probe(0, 20000)
    ffma_outer = probe(1, 20000)
    ffma2_outer = probe(2, 20000)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    bytes_per_step = transition_bytes(D, A)
    state_bytes = 2 * (4 * spec.sf_rows + 8 * spec.sd_rows + 20)      # read + write once per launch
    hbm_gbs = (bytes_per_step * n * K + state_bytes * n) / (kern_ms * 1e-3) / 1e9
    # DRAM traffic of the dominant kernel: ncu --set full capture of this kernel (profiles/r1_tc_rollout_summary.md,
    # profiles/r1_v3_rollout_fused_summary.md: dram read+write = 325 MB / 323 MB for 2^21 env-steps) -> bytes per env-step
    ncu_traffic_per_env_step = {"tc": 155.0, "ffma": 154.0}[args.engine] if args.env == "QuadTracking" else None
    traffic = None if ncu_traffic_per_env_step is None else ncu_traffic_per_env_step * n * K
    tensor_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    tensor_burst = float(peaks.get("bf16_tflops", 1590.0))
    hbm_info = {"achieved_gbs": hbm_gbs, "peak_gbs": hbm_peak, "frac": hbm_gbs / hbm_peak,
                "algorithmic_bytes_per_env_step": bytes_per_step + state_bytes / K,
                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}
    if args.engine == "ffma":
        roofline = {"bound": "fp32_ffma", "achieved": achieved_tflops, "peak": ffma_peak, "unit": "TFLOP/s",
                    "frac": achieved_tflops / ffma_peak, "traffic": traffic,
                    "peak_source": "measured in this run: msacl_ffma_probe (8 independent FFMA chains/thread, 2x256 threads/SM); "
                                   "MEASURED_PEAKS.json has no FP32 figure (theoretical 148*128*2*1.965 GHz = 74.4)",
                    "register_tiled_sgemm_ceiling_tflops": ffma_outer, "register_tiled_ffma2_ceiling_tflops": ffma2_outer,
                    "kernel": f"rollout_fused_kernel<{args.env}>", "kernel_ms": kern_ms,
                    "algorithmic_flops_per_env_step": flops_per_step, "hbm": hbm_info}
    else:
        # tensor path: algorithmic FLOPs (one FP32-equivalent pass) against the measured dense bf16 peak; the
        # split-bf16 scheme issues 3 UMMAs per algorithmic product, so tensor-pipe utilisation is ~3x `frac`.
        roofline = {"bound": "tensor", "achieved": achieved_tflops, "peak": tensor_peak, "unit": "TFLOP/s",
                    "frac": achieved_tflops / tensor_peak, "traffic": traffic,
                    "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum of this kernel, scaled per env-step from the profiled launch",
                    "peak_source": ("MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks
                                    else "fallback 1.4 PFLOP/s sustained"),
                    "peak_burst": tensor_burst, "tensor_issue_factor": 3,
                    "tensor_pipe_frac_incl_split": 3 * achieved_tflops * (1 - 2 * 256 * 2 * A / flops_per_step) / tensor_peak,
                    "fp32_ffma_peak_tflops": ffma_peak,
                    "kernel": f"rollout_tc_kernel<{args.env}>", "kernel_ms": kern_ms,
                    "algorithmic_flops_per_env_step": flops_per_step, "hbm": hbm_info}

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cpu = cpu_baseline(args.env)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.engine == "ffma" else "bf16x3 (split-bf16 tensor-core products, f32 accumulate; f32/f64 dynamics)",
            "data": "synthetic",
            "config": {"workload": f"{args.env} fused rollout (actor MLP + TanhGauss sample + env ODE step + reward/cost + autoreset), "
                                   f"{n} envs/GPU (BASELINE config 5 per-GPU share), {K} env steps per launch, transitions written to HBM",
                       "engine": args.engine, "envs_per_gpu": n, "inner_steps": K, "n_step": n_step, "l2": "inputs larger than L2 (state + transitions per launch >> 126 MB)",
                       "parallelism": f"env-sharded x{world}, no step-path collective"},
            "clocks": clk, "gpu_launches": args.steps,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": e2e_ms / args.steps, "gpu_launches_per_step": 5,
                    "what": "per step: H2D actor weights (pinned) -> sampler rollout -> buffer.add_batch (device window scatter) -> "
                            "buffer.sample_batch -> [one packed NCCL all-gather of the sub-batches + statistics if N>1] -> D2H replay batch + episode stats (pinned)"},
            "roofline": roofline,
        }
        if cpu is not None:
            out["cpu_baseline"] = cpu
        if e2e_full is not None:
            out["e2e_full_transition_d2h"] = e2e_full
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

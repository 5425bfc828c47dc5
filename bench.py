#!/usr/bin/env python
"""bench.py -- env.step()/sec of VSS-v0 3v3 at 65 536 envs per GPU (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]          # this repo's CUDA path
  python bench.py --impl reference ...                        # the CPU path, same metric

One "step" = ONE fused launch advancing all 65 536 matches of one world by one control
step: OU commands + action->wheel conversion, 5 physics sub-steps, observation, reward,
done, truncation, info accumulators, masked auto-reset (rs_vss_env_step).

L2 hygiene: one world's per-step traffic (38.8 MB) fits the 126 MB L2, so the timed loop
rotates over M independent worlds (default 8 -> 310 MB of state+outputs) and each world is
touched again only after 7 other worlds streamed through: every step reads its state from
HBM ("inputs larger than L2").  The K steps are replayed from ONE captured CUDA graph (the
Philox step counter lives in device memory, so replays draw fresh noise).

N > 1: one process per GPU (torchrun), each rank owns its own 65 536-env worlds (global
env ids are disjoint, "weak" scaling), no collective on the step path; the ranks meet only
at the barrier around the timed region and at the max-over-ranks of the device time.
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

ENVS_PER_GPU = 65536
ALG_BYTES_PER_ENV_STEP = 592          # SURVEY.md section 8(d), VSS-v0
METRIC = "env.step()/sec at 65536 VSS-v0 3v3 envs"
UNIT = "env-steps/s"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            c = [x.strip() for x in r.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0])); mx.append(float(c[1]))
            except ValueError:
                continue
            for nme, v in zip(names, c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline_run(threads, target_seconds, envs=4096):
    """The CPU path on the host cores: the oracle's restatement of VSSEnv.step (kind "port";
    robosim itself -- the reference's engine -- cannot be installed)."""
    import numpy as np
    from oracle import oracle as O
    O.build()
    # num_threads() in the oracle's omp pragma overrides OMP_NUM_THREADS (torchrun sets it to 1)
    threads = O.usable_threads(threads)
    w = O.OracleWorld(O.KIND_VSS, 0, 3, 3, 25, envs, seed=1, threads=threads)
    w.task_reset(O.TASK_VSS)
    rng = np.random.default_rng(0)
    act = rng.uniform(-1, 1, (envs, 2)).astype(np.float32)
    for _ in range(3):
        w.vss_env_step(act)
    t0 = time.perf_counter()
    w.vss_env_step(act)
    one = max(time.perf_counter() - t0, 1e-6)
    steps = int(max(5, min(20000, target_seconds / one)))
    t0 = time.perf_counter()
    for _ in range(steps):
        w.vss_env_step(act)
    dt = time.perf_counter() - t0
    return {"value": envs * steps / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d VSS-v0 envs x %d steps of oracle/rs_oracle.c (fp64 C, OpenMP) in %.1f s"
                      % (envs, steps, dt)}


def run_reference(args, rank, world, out):
    """--impl reference: the CPU path, all host threads, bounded sample per step."""
    if rank != 0:
        return
    import numpy as np
    from oracle import oracle as O
    O.build()
    threads = O.usable_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    envs = 8192
    w = O.OracleWorld(O.KIND_VSS, 0, 3, 3, 25, envs, seed=1, threads=threads)
    w.task_reset(O.TASK_VSS)
    rng = np.random.default_rng(0)
    act = rng.uniform(-1, 1, (envs, 2)).astype(np.float32)
    steps = min(args.steps, 400)
    warm = min(args.warmup, 20)
    for _ in range(warm):
        w.vss_env_step(act)
    t0 = time.perf_counter()
    for _ in range(steps):
        w.vss_env_step(act)
    dt = time.perf_counter() - t0
    v = envs * steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "VSS-v0 3v3, 65536 envs per GPU",
                   "note": "robosim (rc-robosim 1.2.0) is not installable here; this arm times the "
                           "CPU restatement of the same path (oracle/rs_oracle.c) on a bounded sample"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d envs x %d steps per run, OpenMP %d threads" % (envs, steps, threads)},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), file=out, flush=True)


def _claim_stdout():
    """stdout must carry exactly ONE JSON line.  Libraries write to file descriptor 1 behind
    Python's back (NCCL prints "NCCL version ..." there at communicator creation), so fd 1 is
    pointed at stderr for the whole run and the JSON line goes to a private copy of the
    original stdout."""
    out = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)
    return out


def main():
    out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=24000, help="timed steps (default = 20 episode horizons)")
    ap.add_argument("--warmup", type=int, default=4800)
    ap.add_argument("--min-warmup", type=int, default=600,
                    help="lower bound on untimed steps PER WORLD (half an episode horizon), so that the timed\n"
                         "region sees the steady-state contact density (robots on walls) and boosted clocks,\n"
                         "not freshly reset scenes (which step ~25 %% faster)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU, help="envs per GPU")
    ap.add_argument("--worlds", type=int, default=8, help="independent worlds rotated through (L2)")
    ap.add_argument("--e2e-steps", type=int, default=100)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world, out)
        return

    import torch
    import torch.distributed as dist
    from rsoccer_b200 import engine as E

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the CUDA path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    N, M = args.envs, max(1, args.worlds)
    W = max(3, args.warmup, args.min_warmup * M)
    K = max(1, args.steps)
    worlds, acts, outs = [], [], []
    gen = torch.Generator(device="cpu").manual_seed(1234 + rank)
    for m in range(M):
        w = E.BatchedWorld(E.KIND_VSS, 0, 3, 3, 25, N, device=dev, seed=2024,
                           env_offset=(rank * M + m) * N)
        w.task_reset(E.TASK_VSS_V0)
        worlds.append(w)
        acts.append((torch.rand(N, 2, generator=gen) * 2 - 1).to(dev))
        outs.append(w.alloc_outputs(E.TASK_VSS_V0))
    torch.cuda.synchronize()
    launches0 = sum(w.launches for w in worlds)

    def step(i):
        m = i % M
        worlds[m].vss_env_step(acts[m], out=outs[m])

    stream = torch.cuda.Stream(device=dev)
    graph = None
    with torch.cuda.stream(stream):
        for i in range(W):
            step(i)
        stream.synchronize()
        if not args.no_graph:
            # one graph = G = 4 passes over the M worlds (-1 % vs one pass per graph launch)
            G = 4 * M
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                for i in range(G):
                    step(i)
            graph.replay()
            stream.synchronize()
    G = 4 * M
    reps, tail = (K // G, K % G) if graph is not None else (0, K)     # EXACTLY K steps are timed

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(reps):
            graph.replay()
        for i in range(tail):
            step(i)
        ev1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    launches = reps * G + tail
    value = N * world * K / (ms * 1e-3)

    # ---- e2e: the public host-buffer call, pinned host memory, H2D + D2H inside the timing
    Ke = max(10, args.e2e_steps)
    w0 = worlds[0]
    h_act = torch.empty(N, 2, dtype=torch.float32).pin_memory()
    h_act.copy_(acts[0].cpu())
    h_obs, h_rew, h_done, h_trunc = w0.alloc_host_outputs(E.TASK_VSS_V0)   # one pinned block -> one D2H copy
    with torch.cuda.stream(stream):
        for _ in range(3):
            w0.vss_env_step_host(h_act, h_obs, h_rew, h_done, h_trunc)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(Ke):
            w0.vss_env_step_host(h_act, h_obs, h_rew, h_done, h_trunc)
        e1.record(stream)
        torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e_value = N * world * Ke / (ms_e2e * 1e-3)

    if rank == 0:
        peak, peak_src = _peaks()
        per_launch_s = ms * 1e-3 / K
        achieved = ALG_BYTES_PER_ENV_STEP * N / per_launch_s / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get("k_vss_env_step_dram_bytes_per_launch")
        except Exception:
            pass
        ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        cpu = cpu_baseline_run(ncpu, args.cpu_seconds)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "VSS-v0 3v3, %d envs per GPU" % N, "envs_per_gpu": N,
                       "field_type": 0, "time_step_ms": 25, "substeps": 5,
                       "l2": "inputs larger than L2: %d independent %d-env worlds rotated (%.0f MB per pass > 126 MB L2)"
                             % (M, N, M * N * ALG_BYTES_PER_ENV_STEP / 1e6),
                       "launch": ("cuda graph replay" if graph is not None else "direct launches")
                                 + ", programmatic dependent launch between consecutive steps",
                       "parallelism": "env-sharded x%d, no collective on the step path" % world},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "steps": Ke, "ms_per_step": ms_e2e / Ke,
                    "h2d_bytes_per_step": N * 2 * 4, "d2h_bytes_per_step": N * (40 * 4 + 4 + 1 + 1),
                    "api": "rs_vss_env_step_host (pinned host buffers: actions read over PCIe by the kernel, one packed D2H of obs/reward/done/trunc, sync)",
                    "pcie_gbs": (N * 2 * 4 + N * (40 * 4 + 4 + 1 + 1)) / (ms_e2e * 1e-3 / Ke) / 1e9},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "k_vss_env_step<3,3,64,2> (VssF0P: compile-time constants, packed fp32x2 forms)",
                         "alg_bytes_per_launch": ALG_BYTES_PER_ENV_STEP * N,
                         "avg_launch_us": per_launch_s * 1e6},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), file=out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- env.step()/sec of the batched robot-soccer engine (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]          # this repo's CUDA path
  python bench.py --impl reference ...                        # the CPU path, same metric
  python bench.py --config {vss65536,vss4096,sd4096,cp16384,vss262144_sharded,vss65536_strong}

One "step" = ONE fused launch advancing every match of one world by one control step:
commands (agent action + OU noise), 5 physics sub-steps, observation, reward, done,
truncation, info accumulators, masked auto-reset (rs_vss_env_step / rs_ssl_env_step).

What is timed.  The K steps of the command line are captured into ONE CUDA graph (the Philox
step counter lives in device memory, so replays draw fresh noise) and the graph is replayed
`repeats` times, until the timed region holds >= --min-ms of device time: a 20-step region
is 0.3 ms, far too short for CUDA events, so ms_per_step = total / (K x repeats).  `steps`,
`repeats`, `warmup` (W direct launches right before the timed region) and `settle_steps` (the
untimed steps that bring every world to its steady-state contact density, robots leaning on
walls: freshly reset scenes step ~25 % faster) are all reported; config.launch says what ran.

L2 hygiene: one world's per-step traffic (38.8 MB at 65 536 matches) fits the 126 MB L2, so
the steps rotate over M independent worlds whose combined traffic exceeds 2.4 x L2, and the
graph length is a multiple of M: every step reads its state from HBM ("inputs larger than
L2").

N > 1: one process per GPU (torchrun), each rank owns its own worlds (global env ids are
disjoint, "weak" scaling), no collective on the step path; the ranks meet only at the barrier
around the timed region and at the max-over-ranks of the device time.  The default line also
carries `configs` (BASELINE configs 2-5 at the sizes BASELINE.json names), `strong` (the
literal "65 536 envs on N GPUs": 65 536 / N per GPU) and, for N > 1, `gather` (the rollout
all-gather over NCCL, alone and overlapped with the next rollout on a side stream).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

UNIT = "env-steps/s"
METRIC = "env.step()/sec at 65536 VSS-v0 3v3 envs"
L2_BYTES = 126e6

# algorithmic bytes per env-step; BASELINE.json configs 2-5
CONFIGS = {
    "vss65536": dict(task="vss", envs=65536, alg=592, max_steps=1200, settle=600,
                     workload="VSS-v0 3v3, 65536 envs per GPU"),
    "vss4096": dict(task="vss", envs=4096, alg=592, max_steps=1200, settle=600,
                    workload="VSS-v0 3v3, 4096 envs (BASELINE config 2)"),
    "sd4096": dict(task="sd", envs=4096, alg=556, max_steps=1000, settle=300,
                   workload="SSLStaticDefenders-v0, 4096 envs (BASELINE config 3)"),
    "cp16384": dict(task="cp", envs=16384, alg=276, max_steps=1200, settle=300,
                    workload="SSLContestedPossession-v0, 16384 envs (BASELINE config 4)"),
    "vss262144_sharded": dict(task="vss", envs=32768, alg=592, max_steps=1200, settle=600,
                              workload="VSS-v0 3v3, 32768 envs per GPU (BASELINE config 5: 262144 over 8 GPUs)"),
    "vss65536_strong": dict(task="vss", envs=None, alg=592, max_steps=1200, settle=600,
                            workload="VSS-v0 3v3, 65536 envs in total, split over the GPUs"),
}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _host_threads():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.times, self.proc = index, [], [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.times.append(time.perf_counter())
            self.rows.append(line.strip())

    def stop(self, t_from=None, t_to=None):
        """statistics of the samples taken in [t_from, t_to] (perf_counter; default: all of them)"""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        if t_from is not None:
            # a sample is printed up to one period (20 ms) after it was taken
            keep = [r for t, r in zip(self.times, self.rows) if t_from <= t <= (t_to if t_to is not None else 1e30) + 0.03]
            if len(keep) >= 2:
                self.rows = keep
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
        pw = []
        for r in self.rows:
            c = [x.strip() for x in r.split(",")]
            try:
                pw.append(float(c[2]))
            except (ValueError, IndexError):
                pass
        lo = min(sm) if sm else None
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "sm_min_mhz": lo, "power_w_max": max(pw) if pw else None}


# ----------------------------------------------------------------------------- CPU arms
def _oracle_world(task, envs, threads):
    from oracle import oracle as O
    O.build()
    threads = O.usable_threads(threads)
    if task == "vss":
        w = O.OracleWorld(O.KIND_VSS, 0, 3, 3, 25, envs, seed=1, threads=threads)
        tid, ad, ms = O.TASK_VSS, 2, 1200
    elif task == "sd":
        w = O.OracleWorld(O.KIND_SSL, 2, 1, 6, 25, envs, seed=1, threads=threads)
        tid, ad, ms = O.TASK_SSL_STATIC_DEFENDERS, 5, 1000
    else:
        w = O.OracleWorld(O.KIND_SSL, 2, 1, 1, 25, envs, seed=1, threads=threads)
        tid, ad, ms = O.TASK_SSL_CONTESTED_POSSESSION, 5, 1200
    w.task_reset(tid)
    import numpy as np
    act = np.random.default_rng(0).uniform(-1, 1, (envs, ad)).astype(np.float32)
    if task == "vss":
        return w, threads, (lambda: w.vss_env_step(act, max_steps=ms))
    return w, threads, (lambda: w.ssl_env_step(tid, act, max_steps=ms))


def cpu_baseline_run(task, threads, target_seconds, envs=4096):
    """The CPU path on the host cores: the oracle's restatement of the env's step (kind "port";
    robosim itself -- the reference's engine -- cannot be installed)."""
    w, threads, step = _oracle_world(task, envs, threads)
    for _ in range(3):
        step()
    t0 = time.perf_counter()
    step()
    one = max(time.perf_counter() - t0, 1e-6)
    steps = int(max(5, min(20000, target_seconds / one)))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return {"value": envs * steps / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d envs x %d steps of oracle/rs_oracle.c (fp64 C, OpenMP) in %.1f s" % (envs, steps, dt)}


def reference_vssenv_run(target_seconds=4.0):
    """the UNMODIFIED reference VSSEnv (its Python wrapper:
    vss_gym.py:89, vss_gym_base.py:72) stepping one env, engine = the oracle behind the robosim
    stand-in.  Needs the reference package installed under baseline/_ref (build() does that when
    /root/reference is present); returns None otherwise."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "rsoccer_gym")):
        return None
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden", "shims"))
        import robosim as oracle_robosim          # tests/golden/shims: `robosim` served by the oracle
        from rsoccer_b200 import compat
        compat.install(robosim_module=oracle_robosim)      # + gymnasium / pygame stand-ins when those are absent
        sys.path.insert(0, ref)
        import gymnasium as gym
        import rsoccer_gym           # noqa: F401  (the unmodified reference package)
        import numpy as np
        env = gym.make("VSS-v0")
        env.reset()
        rng = np.random.default_rng(0)
        for _ in range(50):
            env.step(rng.uniform(-1, 1, 2).astype(np.float32))
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < target_seconds:
            for _ in range(100):
                _, _, done, trunc, _ = env.step(rng.uniform(-1, 1, 2).astype(np.float32))
                if done or trunc:
                    env.reset()
            n += 100
        dt = time.perf_counter() - t0
        return {"value": n / dt, "unit": UNIT, "cores": 1, "us_per_step": 1e6 * dt / n,
                "what": "unmodified reference VSSEnv.step (gym.make('VSS-v0'), 1 env, 1 process), engine = "
                        "oracle/rs_oracle.c behind the robosim stand-in: Python wrapper + physics"}
    except Exception as e:       # the arm must never take the bench down
        return {"unavailable": "%s: %s" % (type(e).__name__, e)}


def run_reference(args, rank, out):
    """--impl reference: the reference's CPU implementation of the path on the host cores.
    robosim (rc-robosim) is tried first; it is absent from this image, so the
    arm times the CPU restatement (oracle/rs_oracle.c, all host threads) at the FULL config size:
    K steps of 65 536 envs, repeated until >= 2 s are timed."""
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    envs = cfg["envs"] or 65536
    K, W = max(1, args.steps), max(0, args.warmup)
    threads = _host_threads()
    try:
        import robosim  # noqa: F401
        have_robosim = True
    except Exception:
        have_robosim = False
    if have_robosim and cfg["task"] == "vss":
        # the real thing: gym.make('VSS-v0') random-action loop, one process (README.md:116-133)
        sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
        import gymnasium as gym
        import numpy as np
        import rsoccer_gym  # noqa: F401
        env = gym.make("VSS-v0")
        env.reset()
        rng = np.random.default_rng(0)
        for _ in range(W):
            env.step(rng.uniform(-1, 1, 2).astype(np.float32))
        reps, t0 = 0, time.perf_counter()
        while reps == 0 or time.perf_counter() - t0 < 2.0:
            for _ in range(K):
                _, _, d, tr, _ = env.step(rng.uniform(-1, 1, 2).astype(np.float32))
                if d or tr:
                    env.reset()
            reps += 1
        dt = time.perf_counter() - t0
        v, kind, threads = K * reps / dt, "reference", 1
        sample = "rSim: gym.make('VSS-v0'), 1 env, 1 process, %d steps in %.1f s" % (K * reps, dt)
        ms_per_step = 1e3 * dt / (K * reps)
    else:
        w, threads, step = _oracle_world(cfg["task"], envs, threads)
        for _ in range(max(W, 1)):
            step()
        reps, t0 = 0, time.perf_counter()
        while reps == 0 or time.perf_counter() - t0 < 2.0:
            for _ in range(K):
                step()
            reps += 1
        dt = time.perf_counter() - t0
        v, kind = envs * K * reps / dt, "port"
        sample = "%d envs x %d steps x %d repeats in %.1f s, OpenMP %d threads" % (envs, K, reps, dt, threads)
        ms_per_step = 1e3 * dt / (K * reps)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": K, "repeats": reps, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["workload"], "envs_per_gpu": envs, "field_type": 0 if cfg["task"] == "vss" else 2,
                   "time_step_ms": 25, "substeps": 5,
                   "robosim": "imported" if have_robosim else "absent (rc-robosim 1.2.0 is not installable here)",
                   "note": "rSim timed through the unmodified reference env" if have_robosim else
                           "this arm times the CPU restatement of the same path (oracle/rs_oracle.c, fp64, OpenMP) "
                           "at the full config size; it is not rSim"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if not have_robosim:
        # the bare C loop on ONE thread next to the all-threads figure
        try:
            one = cpu_baseline_run(cfg["task"], 1, 1.5, envs=4096)
            line["single_thread"] = {"value": one["value"], "unit": UNIT, "cores": one["cores"], "kind": "port", "sample": one["sample"]}
        except Exception as e:
            line["single_thread"] = {"unavailable": "%s: %s" % (type(e).__name__, e)}
    wrapped = reference_vssenv_run(3.0) if cfg["task"] == "vss" and not have_robosim else None
    if wrapped is not None:
        line["reference_vssenv"] = wrapped
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


# ----------------------------------------------------------------------------- host placement
def bind_near_gpu(torch, index):
    """Pin this process to the CPUs of its GPU's NUMA node BEFORE it allocates pinned memory (first
    touch then places the e2e landing zone next to the GPU's PCIe root).  Returns what it found."""
    info = {"numa_node": None, "cpus": None, "bound": False}
    try:
        p = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        with open(base + "/numa_node") as f:
            info["numa_node"] = int(f.read().strip())
        with open(base + "/local_cpulist") as f:
            cl = f.read().strip()
        info["cpus"] = cl
        cpus = set()
        for part in cl.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            info["bound"] = True
        info["n_cpus"] = len(use or allowed)
    except Exception as e:
        info["error"] = "%s: %s" % (type(e).__name__, e)
    return info


# ----------------------------------------------------------------------------- GPU workloads
class Workload:
    """M independent worlds of one config on this rank's GPU, stepped round-robin."""

    def __init__(self, torch, E, name, cfg, envs, dev, rank, overlap, K, seed_base=0, worlds=0):
        from rsoccer_b200 import _lib
        self.torch, self.E, self.name, self.cfg, self.N, self.dev = torch, E, name, cfg, envs, dev
        self.task = cfg["task"]
        # combined per-pass traffic >= 2.4 x L2 (8 worlds at 65 536 matches); the graph holds lcm(K, M) steps, so
        # that a replay continues the rotation where the previous one stopped: M is the first count that keeps it short
        m0 = max(2, int(math.ceil(2.4 * L2_BYTES / (envs * cfg["alg"]))))
        self.M = next((m for m in range(m0, m0 + 4 * K + 1) if K * m // math.gcd(K, m) <= 4096), m0)
        if worlds:
            self.M = worlds          # e.g. 1: one world stepped again and again (state stays in L2; a true dependency chain)
        self.G = K * self.M // math.gcd(K, self.M)
        self.worlds, self.acts, self.outs = [], [], []
        gen = torch.Generator(device="cpu").manual_seed(1234 + rank)
        self.seed_base, self.rank = seed_base, rank
        for m in range(self.M):
            w, ad = self.new_world(envs, ((seed_base * 64 + rank) * self.M + m) * envs)
            w.set_option(_lib.OPT_STEP_OVERLAP, overlap)
            self.worlds.append(w)
            self.acts.append((torch.rand(envs, ad, generator=gen) * 2 - 1).to(dev))
            self.outs.append(w.alloc_outputs(self.tid))
        self.act_dim, self.obs_dim = ad, self.worlds[0].obs_dim(self.tid)
        self.overlap = overlap
        flags = self.worlds[0].kernel_flags
        if self.task == "vss":
            self.kernel = ("k_vss_env_step_lanes<128,F0> (one lane per body)" if flags & 1 else
                           "k_vss_env_step<3,3,64,%d> (one lane per match%s)" % (2 if flags & 8 else 1,
                                                                                  ", packed fp32x2 forms" if flags & 8 else ""))
        else:
            nbny = "1,6" if self.task == "sd" else "1,1"
            self.kernel = ("k_ssl_env_step_lanes<%s> (one lane per body)" % nbny if flags & 1 else
                           "k_ssl_env_step<%s,64> (one lane per match)" % nbny)

    def new_world(self, envs, env_offset):
        E, dev = self.E, self.dev
        if self.task == "vss":
            w = E.BatchedWorld(E.KIND_VSS, 0, 3, 3, 25, envs, device=dev, seed=2024, env_offset=env_offset)
            self.tid, ad = E.TASK_VSS_V0, 2
        elif self.task == "sd":
            w = E.BatchedWorld(E.KIND_SSL, 2, 1, 6, 25, envs, device=dev, seed=2024, env_offset=env_offset)
            self.tid, ad = E.TASK_SSL_STATIC_DEFENDERS_V0, 5
        else:
            w = E.BatchedWorld(E.KIND_SSL, 2, 1, 1, 25, envs, device=dev, seed=2024, env_offset=env_offset)
            self.tid, ad = E.TASK_SSL_CONTESTED_POSSESSION_V0, 5
        w.task_reset(self.tid)
        return w, ad

    def step(self, i):
        m = i % self.M
        if self.task == "vss":
            self.worlds[m].vss_env_step(self.acts[m], out=self.outs[m], max_steps=self.cfg["max_steps"])
        else:
            self.worlds[m].ssl_env_step(self.tid, self.acts[m], out=self.outs[m], max_steps=self.cfg["max_steps"])

    def set_overlap(self, mode):
        from rsoccer_b200 import _lib
        for w in self.worlds:
            w.set_option(_lib.OPT_STEP_OVERLAP, mode)
        self.overlap = mode

    def capture(self, stream, n_streams=1):
        """One CUDA graph of G steps.  n_streams > 1: the worlds are spread over that many side streams, so the
        chains of different worlds are PARALLEL branches of the graph (each world's own steps stay a chain on its
        stream) -- how independent small worlds are driven when launch latency, not the GPU, is the limit."""
        torch = self.torch
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            if n_streams <= 1:
                for i in range(self.G):
                    self.step(i)
            else:
                side = [torch.cuda.Stream(device=self.dev) for _ in range(n_streams)]
                for sd in side:
                    sd.wait_stream(stream)
                for i in range(self.G):
                    with torch.cuda.stream(side[(i % self.M) % n_streams]):
                        self.step(i)
                for sd in side:
                    stream.wait_stream(sd)
        return g, self.G

    def host_step(self, bufs):
        w0 = self.worlds[0]
        if self.task == "vss":
            w0.vss_env_step_host(*bufs, max_steps=self.cfg["max_steps"])
        else:
            w0.ssl_env_step_host(self.tid, *bufs, max_steps=self.cfg["max_steps"])

    def close(self):
        for w in self.worlds:
            w.close()
        self.worlds = []


def measure(torch, dist, wl, K, W, min_ms, stream, world, dev, settle=True, e2e_steps=50, clock_index=None, pipelined=False,
            n_streams=1):
    """settle -> capture -> W warm-up launches -> timed graph replays; then the e2e loop."""
    res = {}
    sampler = None
    if clock_index is not None:
        sampler = ClockSampler(clock_index)      # started here, long before the timed region: no idle pause in front of it
        sampler.start()
    with torch.cuda.stream(stream):
        if settle:
            for i in range(wl.cfg["settle"] * wl.M):
                wl.step(i)
        stream.synchronize()
        graph, G = wl.capture(stream, n_streams)
        graph.replay()
        for i in range(W):
            wl.step(i)
        # size the timed region
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); graph.replay(); e1.record(stream)
        stream.synchronize()
        one = max(e0.elapsed_time(e1), 1e-3)
    replays = max(1, int(math.ceil(min_ms / one)))
    if world > 1:
        t = torch.tensor([replays], device=dev, dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        replays = int(t.item())
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_from = time.perf_counter()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(replays):
            graph.replay()
        ev1.record(stream)
    torch.cuda.synchronize()
    t_to = time.perf_counter()
    if world > 1:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    if sampler is not None:
        res["clocks"] = sampler.stop(t_from, t_to)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    steps_timed = replays * G
    res.update(ms_total=ms, steps_timed=steps_timed, graph_len=G, replays=replays,
               ms_per_step=ms / steps_timed, repeats=steps_timed / K)

    # ---- e2e: the public host-buffer call, pinned host memory, H2D + D2H inside the timing
    if e2e_steps > 0:
        w0 = wl.worlds[0]
        N = wl.N
        h_act = torch.empty(N, wl.act_dim, dtype=torch.float32).pin_memory()
        h_act.copy_(wl.acts[0].cpu())
        h_out = w0.alloc_host_outputs(wl.tid)                   # one pinned block -> one D2H copy
        bufs = (h_act,) + tuple(h_out)
        with torch.cuda.stream(stream):
            for _ in range(3):
                wl.host_step(bufs)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(e2e_steps):
                wl.host_step(bufs)
            e1.record(stream)
            torch.cuda.synchronize()
        ms_e = e0.elapsed_time(e1)
        # the box's own ceiling for this transfer: plain D2H copies of the same block, all ranks at once
        d2h = N * (wl.obs_dim * 4 + 4 + 1 + 1)
        dsrc = torch.empty(d2h, dtype=torch.uint8, device=dev)
        hdst = torch.empty(d2h, dtype=torch.uint8).pin_memory()
        with torch.cuda.stream(stream):
            for _ in range(3):
                hdst.copy_(dsrc, non_blocking=True)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record(stream)
            for _ in range(e2e_steps):
                hdst.copy_(dsrc, non_blocking=True)
                stream.synchronize()
            c1.record(stream)
            torch.cuda.synchronize()
        ms_c = c0.elapsed_time(c1)
        if world > 1:
            t = torch.tensor([ms_e, ms_c], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e, ms_c = float(t[0].item()), float(t[1].item())
        # the same through the split-phase calls: TWO env groups of N / 2 matches on two streams, group A steps while
        # group B's outputs cross PCIe (each group: wait for its outputs, start its next step, as a consumer that
        # works on one group at a time would).  Every step still moves its actions in and its outputs out.
        pipe = None
        if pipelined and N % 2 == 0:
            n2 = N // 2
            base = ((wl.seed_base * 64 + wl.rank + 32) * 64) * N
            groups, gbufs, gst = [], [], [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
            for k in range(2):
                wk, _ = wl.new_world(n2, base + k * n2)
                ha = torch.empty(n2, wl.act_dim, dtype=torch.float32).pin_memory()
                ha.copy_(wl.acts[0][k * n2:(k + 1) * n2].cpu())
                groups.append(wk)
                gbufs.append((ha,) + tuple(wk.alloc_host_outputs(wl.tid)))
                da = wl.acts[0][k * n2:(k + 1) * n2].contiguous()
                for _ in range(wl.cfg["settle"]):             # the contact density of the timed worlds
                    if wl.task == "vss":
                        wk.vss_env_step(da, max_steps=wl.cfg["max_steps"])
                    else:
                        wk.ssl_env_step(wl.tid, da, max_steps=wl.cfg["max_steps"])
            torch.cuda.synchronize()

            def begin(k):
                with torch.cuda.stream(gst[k]):
                    if wl.task == "vss":
                        groups[k].vss_env_step_host_begin(*gbufs[k], max_steps=wl.cfg["max_steps"])
                    else:
                        groups[k].ssl_env_step_host_begin(wl.tid, *gbufs[k], max_steps=wl.cfg["max_steps"])

            def run(steps):
                begin(0); begin(1)
                for _ in range(steps - 1):
                    groups[0].host_step_wait(); begin(0)
                    groups[1].host_step_wait(); begin(1)
                groups[0].host_step_wait(); groups[1].host_step_wait()

            run(3)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            p0 = torch.cuda.Event(enable_timing=True)
            pe = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)]
            p0.record(gst[0])                                  # before anything of either group is enqueued
            run(e2e_steps)
            pe[0].record(gst[0]); pe[1].record(gst[1])
            torch.cuda.synchronize()
            ms_p = max(p0.elapsed_time(pe[0]), p0.elapsed_time(pe[1]))
            if world > 1:
                t = torch.tensor([ms_p], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms_p = float(t.item())
            for wk in groups:
                wk.close()
            pipe = {"value": N * world * e2e_steps / (ms_p * 1e-3), "unit": UNIT, "steps": e2e_steps,
                    "ms_per_step": ms_p / e2e_steps, "frac_of_ceiling": ms_c / ms_p,
                    "what": "two env groups of %d matches on two streams through rs_*_env_step_host_begin / "
                            "rs_host_step_wait: one group steps while the other's outputs cross PCIe; CUDA events, "
                            "first enqueue to the later of the two streams' ends; same bytes per step as the blocking call" % n2}
        h2d = N * wl.act_dim * 4
        res["e2e"] = {
            "value": N * world * e2e_steps / (ms_e * 1e-3), "unit": UNIT, "steps": e2e_steps,
            "ms_per_step": ms_e / e2e_steps, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "api": ("rs_vss_env_step_host (pinned host buffers: actions read over PCIe by the kernel" if wl.task == "vss" else
                    "rs_ssl_env_step_host (pinned host buffers: actions staged with one H2D copy") +
                   ", one packed D2H of obs/reward/done/trunc, sync)",
            "pcie_gbs": (h2d + d2h) / (ms_e * 1e-3 / e2e_steps) / 1e9,
            "d2h_ceiling_gbs": d2h / (ms_c * 1e-3 / e2e_steps) / 1e9,
            "d2h_ceiling_what": "plain cudaMemcpyAsync of the same %d-byte block + sync, %d rank(s) at once: the box's own "
                                "limit for this transfer" % (d2h, world),
            "frac_of_ceiling": (ms_c / ms_e),
        }
        if pipe:
            res["e2e"]["pipelined"] = pipe
    del graph
    return res


def small_world_regimes(torch, dist, E, wl, rec, name, c, envs, K, W, args, stream, world, dev, rank, peak, seed_base):
    """Small worlds are bound by launch latency, not by the GPU; two more regimes for the record (wl is closed,
    the single-world workload is returned for the caller to close):
    (1) the worlds of the rotation as PARALLEL branches of the graph (16 streams);
    (2) ONE world stepped again and again: what a user of exactly this config sees per step."""
    rp = measure(torch, dist, wl, K, W, args.min_ms / 2, stream, world, dev, settle=False, e2e_steps=0, n_streams=16)
    rec["parallel_streams"] = {
        "streams": 16, "ms_per_step": rp["ms_per_step"], "value": envs * world / (rp["ms_per_step"] * 1e-3),
        "frac": c["alg"] * envs / (rp["ms_per_step"] * 1e-3) / 1e9 / peak,
        "what": "the same worlds captured on 16 streams: the chains of different worlds are parallel branches "
                "of the graph instead of one chain of programmatic launches"}
    wl.close()
    w1 = Workload(torch, E, name, c, envs, dev, rank, 2, K, seed_base=seed_base, worlds=1)
    rc1 = measure(torch, dist, w1, K, W, args.min_ms / 4, stream, world, dev, e2e_steps=0)
    w1.set_overlap(0)
    rs1 = measure(torch, dist, w1, K, W, args.min_ms / 4, stream, world, dev, settle=False, e2e_steps=0)
    rec["single_world"] = {
        "overlap2_ms_per_step": rc1["ms_per_step"], "serialized_ms_per_step": rs1["ms_per_step"],
        "overlap2_value": envs * world / (rc1["ms_per_step"] * 1e-3),
        "serialized_value": envs * world / (rs1["ms_per_step"] * 1e-3),
        "what": "one world of this size per GPU stepped back to back (a dependency chain; launch / latency bound)"}
    return w1


def sustained_run(torch, dist, wl, stream, world, dev, seconds, clock_index):
    """The same graph replayed for `seconds` (default 2 s) instead of the ~60 ms of the headline region: under
    sustained load the overlapped step kernel draws the board's power limit and the SM clock settles below its
    boost value (sw_power_cap), so burst and sustained throughput differ; both are reported."""
    with torch.cuda.stream(stream):
        graph, G = wl.capture(stream)
        graph.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); graph.replay(); e1.record(stream)
        stream.synchronize()
        one = max(e0.elapsed_time(e1), 1e-3)
    replays = max(1, int(math.ceil(seconds * 1e3 / one)))
    if world > 1:
        t = torch.tensor([replays], device=dev, dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        replays = int(t.item())
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(clock_index)
    sampler.start()
    # the second half is timed: by then power and clocks have settled (they take ~1 s)
    half = replays // 2
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    with torch.cuda.stream(stream):
        ev[0].record(stream)
        for _ in range(half):
            graph.replay()
        ev[1].record(stream)
        for _ in range(replays - half):
            graph.replay()
        ev[2].record(stream)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms_all, ms_tail = ev[0].elapsed_time(ev[2]), ev[1].elapsed_time(ev[2])
    if world > 1:
        t = torch.tensor([ms_all, ms_tail], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_all, ms_tail = float(t[0].item()), float(t[1].item())
    del graph
    per = ms_tail / ((replays - half) * G)
    return {"ms_per_step": per, "value": wl.N * world / (per * 1e-3), "timed_region_ms": ms_tail,
            "whole_run_ms": ms_all, "whole_run_ms_per_step": ms_all / (replays * G),
            "frac": wl.cfg["alg"] * wl.N / (per * 1e-3) / 1e9 / _peaks()[0], "clocks": clocks,
            "what": "the headline graph replayed for %.1f s; the second half is timed (power and clocks settle within ~1 s); "
                    "clocks / power / throttle reasons sampled over the whole run" % (ms_all * 1e-3)}


def summarise(wl, res, K, W, world, peak, peak_src):
    per_launch_s = res["ms_per_step"] * 1e-3
    achieved = wl.cfg["alg"] * wl.N / per_launch_s / 1e9
    d = {
        "workload": wl.cfg["workload"], "envs_per_gpu": wl.N, "value": wl.N * world / per_launch_s, "unit": UNIT,
        "ms_per_step": res["ms_per_step"], "steps": K, "repeats": res["repeats"], "warmup": W,
        "graph_len": res["graph_len"], "worlds_rotated": wl.M, "settle_steps": wl.cfg["settle"] * wl.M,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "alg_bytes_per_env_step": wl.cfg["alg"], "alg_bytes_per_launch": wl.cfg["alg"] * wl.N,
                     "avg_launch_us": per_launch_s * 1e6, "kernel": wl.kernel, "peak_source": peak_src},
    }
    if "e2e" in res:
        d["e2e"] = res["e2e"]
    return d


def gather_bench(torch, dist, wl, stream, dev, world, T=8):
    """the one collective of the design -- the all-gather that concatenates rollout
    tensors -- alone, and on a side stream while the next rollout runs (rsoccer_b200.sharding)."""
    from rsoccer_b200.sharding import gather_rollout
    N, D = wl.N, wl.obs_dim
    traj = torch.empty(T, N, D, dtype=torch.float32, device=dev)
    side = torch.cuda.Stream(device=dev)

    def rollout():
        for i in range(T):
            wl.step(i)
            traj[i].copy_(wl.outs[i % wl.M][0])

    def timed(fn, reps=5):
        fn()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    with torch.cuda.stream(stream):
        ms_roll = timed(rollout)
        ms_gath = timed(lambda: gather_rollout(traj, n_total=N * world, dim=1))

        # the next rollout overwrites traj while the gather reads it: double-buffer as a consumer would
        traj2 = traj.clone()

        def both_db():
            out, ev = gather_rollout(traj2, n_total=N * world, dim=1, stream=side)
            rollout()
            stream.wait_event(ev)
            return out
        ms_both = timed(both_db)
    recv = T * N * D * 4 * world
    return {"what": "all_gather_into_tensor of obs[T=%d, %d, %d] f32 per rank over NCCL (rsoccer_b200.sharding.gather_rollout)" % (T, N, D),
            "bytes_out_per_rank": recv, "ms_gather": ms_gath, "algbw_gbs": recv / (ms_gath * 1e-3) / 1e9,
            "busbw_gbs": recv * (world - 1) / world / (ms_gath * 1e-3) / 1e9,
            "ms_rollout_alone": ms_roll, "ms_rollout_with_gather_on_side_stream": ms_both,
            "overlap_hidden_frac": max(0.0, min(1.0, (ms_roll + ms_gath - ms_both) / ms_gath))}


def main():
    out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=240, help="K: steps per timed group (one CUDA graph holds a multiple of K)")
    ap.add_argument("--warmup", type=int, default=8, help="W: direct launches right before the timed region (>= 3)")
    ap.add_argument("--min-ms", type=float, default=60.0, help="the timed region replays the graph until it holds this much device time")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="vss65536", choices=sorted(CONFIGS))
    ap.add_argument("--envs", type=int, default=0, help="override envs per GPU of the main config")
    ap.add_argument("--overlap", type=int, default=int(os.environ.get("RS_BENCH_OVERLAP", "3")), choices=[0, 1, 2, 3],
                    help="RS_OPT_STEP_OVERLAP of the timed worlds (include/rsoccer_b200.h)")
    ap.add_argument("--e2e-steps", type=int, default=100)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--sustained-seconds", type=float, default=2.0, help="length of the `sustained` run (0 = skip)")
    ap.add_argument("--no-extras", action="store_true", help="main config only: no configs / strong / gather / serialized sub-records")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, out)
        return

    import torch
    import torch.distributed as dist
    from rsoccer_b200 import engine as E

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the CUDA path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    placement = bind_near_gpu(torch, local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    K, W = max(1, args.steps), max(3, args.warmup)
    peak, peak_src = _peaks()
    stream = torch.cuda.Stream(device=dev)
    cfg = dict(CONFIGS[args.config])
    envs = args.envs or cfg["envs"] or max(1, 65536 // world)

    wl = Workload(torch, E, args.config, cfg, envs, dev, rank, args.overlap, K)
    launches0 = sum(w.launches for w in wl.worlds)
    res = measure(torch, dist, wl, K, W, args.min_ms, stream, world, dev, e2e_steps=max(10, args.e2e_steps),
                  clock_index=local_rank, pipelined=True)
    main_sum = summarise(wl, res, K, W, world, peak, peak_src)
    launches_timed = res["steps_timed"]
    extras = {}
    if args.sustained_seconds > 0:
        extras["sustained"] = sustained_run(torch, dist, wl, stream, world, dev, args.sustained_seconds, local_rank)
    if not args.no_extras:
        # the same worlds with the grid-wide wait between steps (RS_OPT_STEP_OVERLAP = 0), for the record
        if args.overlap != 0:
            wl.set_overlap(0)
            r0 = measure(torch, dist, wl, K, W, args.min_ms / 2, stream, world, dev, settle=False, e2e_steps=0)
            extras["serialized"] = {"ms_per_step": r0["ms_per_step"], "value": envs * world / (r0["ms_per_step"] * 1e-3),
                                    "frac": cfg["alg"] * envs / (r0["ms_per_step"] * 1e-3) / 1e9 / peak,
                                    "what": "same graph with RS_OPT_STEP_OVERLAP = 0: every step waits for the whole previous grid"}
            wl.set_overlap(args.overlap)
        if world > 1 and cfg["task"] == "vss":
            extras["gather"] = gather_bench(torch, dist, wl, stream, dev, world)
        if cfg["task"] == "vss":
            # ONE world stepped again and again: step k+1 of a tile really waits for step k of that tile
            # (state L2 resident -- this is a latency figure, not an HBM one; mode 2: the 7-CTA build)
            w1 = Workload(torch, E, args.config, cfg, envs, dev, rank, 2, K, seed_base=9, worlds=1)
            rc = measure(torch, dist, w1, K, W, args.min_ms / 2, stream, world, dev, e2e_steps=0)
            w1.set_overlap(0)
            rs = measure(torch, dist, w1, K, W, args.min_ms / 2, stream, world, dev, settle=False, e2e_steps=0)
            extras["single_world"] = {
                "what": "one %d-env world stepped back to back (no rotation: state stays in the 126 MB L2, every step "
                        "depends on the previous one)" % envs,
                "overlap2_ms_per_step": rc["ms_per_step"], "serialized_ms_per_step": rs["ms_per_step"],
                "overlap2_value": envs * world / (rc["ms_per_step"] * 1e-3),
                "serialized_value": envs * world / (rs["ms_per_step"] * 1e-3)}
            w1.close()
            del w1
    wl.close()
    del wl
    torch.cuda.empty_cache()

    if not args.no_extras:
        subs = {}
        names = [n for n in ("vss4096", "sd4096", "cp16384", "vss262144_sharded") if n != args.config]
        for name in names:
            c = dict(CONFIGS[name])
            w2 = Workload(torch, E, name, c, c["envs"], dev, rank, args.overlap, K, seed_base=1 + names.index(name))
            r2 = measure(torch, dist, w2, K, W, args.min_ms / 2, stream, world, dev, e2e_steps=30)
            subs[name] = summarise(w2, r2, K, W, world, peak, peak_src)
            if c["envs"] <= 16384:
                w2 = small_world_regimes(torch, dist, E, w2, subs[name], name, c, c["envs"], K, W, args, stream, world, dev, rank,
                                         peak, 11 + names.index(name))
            w2.close()
            del w2
            torch.cuda.empty_cache()
        extras["configs"] = subs
        if args.config != "vss65536_strong":
            c = dict(CONFIGS["vss65536_strong"])
            n_s = max(1, 65536 // world)
            w3 = Workload(torch, E, "vss65536_strong", c, n_s, dev, rank, args.overlap, K, seed_base=7)
            r3 = measure(torch, dist, w3, K, W, args.min_ms / 2, stream, world, dev, e2e_steps=30)
            extras["strong"] = summarise(w3, r3, K, W, world, peak, peak_src)
            extras["strong"]["scaling"] = "strong"
            extras["strong"]["envs_total"] = n_s * world
            if n_s <= 16384:
                w3 = small_world_regimes(torch, dist, E, w3, extras["strong"], "vss65536_strong", c, n_s, K, W, args, stream, world,
                                         dev, rank, peak, 15)
            w3.close()
            del w3

    if rank == 0:
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
            traffic = tj.get("k_vss_env_step_dram_bytes_per_launch") if cfg["task"] == "vss" and envs == 65536 else None
            traffic_src = tj.get("source")
        except Exception:
            pass
        cpu = cpu_baseline_run(cfg["task"], _host_threads(), args.cpu_seconds)
        modes = {0: "every step begins with a grid-wide wait on its predecessor (programmatic dependent launch hides the launch latency only)",
                 1: "RS_OPT_STEP_OVERLAP=1: steps synchronise per 32-match tile on the world state and grid-wide before reading the action buffer",
                 2: "RS_OPT_STEP_OVERLAP=2: steps synchronise per 32-match tile only (fixed action buffers)",
                 3: "RS_OPT_STEP_OVERLAP=3: steps synchronise per 32-match tile only (fixed action buffers) and the step kernel is "
                    "built for 11 resident CTAs per SM; consecutive launches belong to different worlds of the rotation, so they "
                    "are independent and overlap freely -- see `serialized` (grid-wide wait between steps) and `single_world` "
                    "(one world stepped again and again: a true dependency chain) for the other two regimes"}
        line = {
            "metric": METRIC, "value": main_sum["value"], "unit": UNIT, "n_gpus": world, "steps": K,
            "repeats": main_sum["repeats"], "warmup": W, "ms_per_step": main_sum["ms_per_step"],
            "higher_is_better": True, "scaling": "strong" if args.config == "vss65536_strong" else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["workload"], "envs_per_gpu": envs, "field_type": 0 if cfg["task"] == "vss" else 2,
                       "time_step_ms": 25, "substeps": 5,
                       "l2": "inputs larger than L2: %d independent %d-env worlds rotated (%.0f MB per pass > 126 MB L2)"
                             % (main_sum["worlds_rotated"], envs, main_sum["worlds_rotated"] * envs * cfg["alg"] / 1e6),
                       "launch": "CUDA graph of %d steps (%d x K) replayed %d times = %d timed launches (a ~60 ms region at boost "
                                 "clocks; `sustained` is the same graph over 2 s, where the board's power limit lowers the SM clock); %s"
                                 % (res["graph_len"], res["graph_len"] // K, res["replays"], res["steps_timed"], modes[args.overlap]),
                       "settle_steps": main_sum["settle_steps"],
                       "timed_region_ms": res["ms_total"],
                       "parallelism": "env-sharded x%d, no collective on the step path" % world,
                       "host_placement": placement},
            "clocks": res.get("clocks"),
            "e2e": res["e2e"],
            "gpu_launches": launches_timed,
            "roofline": dict(main_sum["roofline"], traffic=traffic,
                             traffic_source=traffic_src if traffic is not None else None),
            "cpu_baseline": cpu,
        }
        line.update(extras)
        print(json.dumps(line), file=out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

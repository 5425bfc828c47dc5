#!/usr/bin/env python
"""Burst vs sustained: the benchmark rotation (8 x 65 536 VSS-v0 matches, graphs of 40 steps) run for --seconds, us per
step per 100 ms window next to nvidia-smi's SM clock / power / throttle reasons sampled every 20 ms."""
import argparse
import os
import subprocess
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rsoccer_b200 import _lib, engine as E  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=4.0)
ap.add_argument("--mode", type=int, default=3)
ap.add_argument("--idle", type=float, default=0.25, help="idle pause before the run (bench.py sleeps 0.25 s for its clock sampler)")
a = ap.parse_args()
n_worlds, envs, glen = 8, 65536, 40
dev = torch.device("cuda", 0)
gen = torch.Generator().manual_seed(1234)
worlds, acts, outs = [], [], []
for m in range(n_worlds):
    w = E.BatchedWorld(E.KIND_VSS, 0, 3, 3, 25, envs, device=dev, seed=2024, env_offset=m * envs)
    w.set_option(_lib.OPT_STEP_OVERLAP, a.mode)
    w.task_reset(E.TASK_VSS_V0)
    worlds.append(w)
    acts.append((torch.rand(envs, 2, generator=gen) * 2 - 1).to(dev))
    outs.append(w.alloc_outputs(E.TASK_VSS_V0))
rows = []
Q = "clocks.sm,clocks.max.sm,power.draw,temperature.gpu,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown"
proc = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=" + Q, "--format=csv,noheader,nounits", "-lms", "20"],
                        stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
threading.Thread(target=lambda: [rows.append((time.perf_counter(), ln.strip())) for ln in proc.stdout], daemon=True).start()
st = torch.cuda.Stream(device=dev)
with torch.cuda.stream(st):
    for i in range(600 * n_worlds):
        worlds[i % n_worlds].vss_env_step(acts[i % n_worlds], out=outs[i % n_worlds])
    st.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
        for i in range(glen):
            worlds[i % n_worlds].vss_env_step(acts[i % n_worlds], out=outs[i % n_worlds])
    g.replay()
    st.synchronize()
    time.sleep(a.idle)
    t_start = time.perf_counter()
    win = []
    while time.perf_counter() - t_start < a.seconds:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(st)
        for _ in range(280):                 # ~100 ms
            g.replay()
        e1.record(st)
        st.synchronize()
        win.append((t0 - t_start, time.perf_counter() - t_start, e0.elapsed_time(e1) * 1e3 / (280 * glen)))
proc.terminate()
for t0, t1, us in win:
    smp = [r for (t, r) in rows if t0 <= t - t_start <= t1]
    clk = [float(r.split(",")[0]) for r in smp if r.split(",")[0].strip().replace(".", "").isdigit()]
    pw = [float(r.split(",")[2]) for r in smp if len(r.split(",")) > 2]
    act = lambda v: v.strip().lower().startswith("active")     # "Active" / "Not Active"
    flags = sorted({" ".join(n for n, v in zip(("sw_power_cap", "hw_slowdown", "sw_thermal_slowdown"), r.split(",")[4:7]) if act(v))
                    for r in smp if len(r.split(",")) > 6})
    print("SUST t=%5.2f..%5.2f s  %.2f us/step   sm %s MHz  power %s W  %s" % (
        t0, t1, us, ("%.0f-%.0f" % (min(clk), max(clk))) if clk else "?", ("%.0f-%.0f" % (min(pw), max(pw))) if pw else "?", ",".join(f for f in flags if f)), flush=True)

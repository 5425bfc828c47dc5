#!/usr/bin/env python
"""Launch floor of a fused step: a CUDA graph of N launches of an EMPTY kernel with the grid, CTA size, parameter
block and launch attributes of a lane-per-match step (rs_debug_empty_step), timed like bench.py times real steps.

  python tools/launch_floor.py [--envs 4096,16384,65536] [--steps 2000]

chain 0 = every launch waits for its predecessor (griddepcontrol.wait), chain 2 = it only triggers its dependents
(what an overlapped step does).  Prints us per launch; the small BASELINE configs are measured against these."""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rsoccer_b200 import _lib, engine as E  # noqa: E402


def floor(envs, chain, steps, glen=200, worlds=8):
    ws = [E.BatchedWorld(E.KIND_VSS, 0, 3, 3, 25, envs, seed=1, env_offset=i * envs) for i in range(worlds)]
    s = torch.cuda.Stream()
    L = _lib.lib()

    def launch(i):
        w = ws[i % worlds]
        _lib.check(L.rs_debug_empty_step(w.h, chain, C.c_void_p(s.cuda_stream)), "rs_debug_empty_step")
    with torch.cuda.stream(s):
        for i in range(50):
            launch(i)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(glen):
                launch(i)
        g.replay()
        s.synchronize()
        reps = max(1, steps // glen)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(reps):
            g.replay()
        e1.record(s)
        s.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * glen)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", default="4096,16384,65536")
    ap.add_argument("--steps", type=int, default=4000)
    a = ap.parse_args()
    for n in [int(x) for x in a.envs.split(",")]:
        print("FLOOR envs=%d  serialised %.2f us/launch   overlapped %.2f us/launch   (graph of 200 empty launches, grid %d x 64)"
              % (n, floor(n, 0, a.steps), floor(n, 2, a.steps), (n + 63) // 64))

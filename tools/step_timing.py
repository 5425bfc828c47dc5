#!/usr/bin/env python
"""Device time per fused env.step() launch, any task / world size / kernel mapping.

  python tools/step_timing.py --task vss|sd|cp|step_vss|step_ssl --envs N [--worlds M] [--steps K]

Same method as bench.py (M independent worlds rotated so that state comes from HBM, one
captured CUDA graph per pass, CUDA events on the launching stream, warm-up past the
steady-state contact density) without the e2e / CPU legs.  Kernel choice and tuning come
from the environment: RS_PER_MATCH=0|1, RS_BLOCK, RS_LANE_BLOCK, RS_PDL=0|1, RS_LIB.
Prints one line: task envs mode us_per_step Menv_steps_per_s.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

from rsoccer_b200 import engine as E  # noqa: E402

AUTO_RESET = os.environ.get("RS_NO_AUTORESET", "") != "1"

TASKS = {
    "vss": (E.KIND_VSS, 0, 3, 3, E.TASK_VSS_V0, 2),
    "sd": (E.KIND_SSL, 2, 1, 6, E.TASK_SSL_STATIC_DEFENDERS_V0, 5),
    "cp": (E.KIND_SSL, 2, 1, 1, E.TASK_SSL_CONTESTED_POSSESSION_V0, 5),
    "step_vss": (E.KIND_VSS, 0, 3, 3, None, 0),
    "step_vss5": (E.KIND_VSS, 1, 5, 5, None, 0),
    "step_vss21": (E.KIND_VSS, 0, 2, 1, None, 0),
    "step_ssl": (E.KIND_SSL, 2, 1, 6, None, 0),
    "step_ssl11": (E.KIND_SSL, 1, 11, 11, None, 0),
}


def time_steps(task_name, envs, n_worlds=8, steps=4000, warmup=300, use_graph=True, graph_steps=0, n_streams=1):
    """us per launch of one fused step of `task_name` at `envs` matches (see module docstring)."""
    kind, ft, nb, ny, task, adim = TASKS[task_name]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    gen = torch.Generator(device="cpu").manual_seed(7)
    worlds, acts, outs = [], [], []
    for m in range(n_worlds):
        w = E.BatchedWorld(kind, ft, nb, ny, 25, envs, device=dev, seed=11, env_offset=m * envs)
        if task is not None:
            w.task_reset(task)
            acts.append((torch.rand(envs, adim, generator=gen) * 2 - 1).to(dev))
            outs.append(w.alloc_outputs(task))
        else:
            w.task_reset(E.TASK_VSS_V0) if (kind == E.KIND_VSS and nb == 3 and ny == 3) else None
            c = torch.rand(envs, nb + ny, w.cmd_dim, generator=gen) * 2 - 1
            if kind == E.KIND_VSS:
                c = c * 40.0
            else:
                c[..., 0] = 0.0
                c[..., 4:] = 0.0
            acts.append(c.to(dev))
            outs.append(None)
        worlds.append(w)

    def step(i):
        m = i % n_worlds
        if task is None:
            worlds[m].step(acts[m])
        elif task == E.TASK_VSS_V0:
            worlds[m].vss_env_step(acts[m], out=outs[m], auto_reset=AUTO_RESET)
        else:
            worlds[m].ssl_env_step(task, acts[m], out=outs[m], auto_reset=AUTO_RESET)

    M = max(n_worlds, graph_steps)          # steps per captured graph (a multiple of the world count)
    stream = torch.cuda.Stream(device=dev)
    graph = None
    with torch.cuda.stream(stream):
        for i in range(warmup * M):
            step(i)
        stream.synchronize()
        if use_graph:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                if n_streams <= 1:
                    for i in range(M):
                        step(i)
                else:
                    # one stream per group of worlds: the chains of different worlds are parallel branches of the
                    # graph, each world's own steps stay a chain (programmatic edges) on its stream
                    side = [torch.cuda.Stream(device=dev) for _ in range(n_streams)]
                    for sd in side:
                        sd.wait_stream(stream)
                    for i in range(M):
                        with torch.cuda.stream(side[(i % n_worlds) % n_streams]):
                            step(i)
                    for sd in side:
                        stream.wait_stream(sd)
            graph.replay()
            stream.synchronize()
        reps = max(1, steps // M)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for r in range(reps):
            if graph is not None:
                graph.replay()
            else:
                for i in range(M):
                    step(i)
        e1.record(stream)
        stream.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * M)
    for w in worlds:
        w.close()
    return us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", default="vss", choices=sorted(TASKS))
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--worlds", type=int, default=8)
    ap.add_argument("--steps", type=int, default=4000)
    ap.add_argument("--warmup", type=int, default=300, help="untimed steps per world")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--graph-steps", type=int, default=0, help="steps per captured graph (default: one pass over the worlds)")
    ap.add_argument("--streams", type=int, default=1, help="capture the worlds on this many streams (parallel branches of the graph)")
    a = ap.parse_args()
    us = time_steps(a.task, a.envs, a.worlds, a.steps, a.warmup, not a.no_graph, a.graph_steps, a.streams)
    mode = "per_match" if os.environ.get("RS_PER_MATCH", "") == "1" else ("per_body" if os.environ.get("RS_PER_MATCH", "") == "0" else "auto")
    print("TIMING task=%s envs=%d mode=%s pdl=%s graph=%d  %.2f us/step  %.1f Menv-steps/s" % (
        a.task, a.envs, mode, os.environ.get("RS_PDL", "1"), not a.no_graph, us, a.envs / us))


if __name__ == "__main__":
    main()

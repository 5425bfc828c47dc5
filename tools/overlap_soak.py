#!/usr/bin/env python
"""Soak of the step-overlap protocol (RS_OPT_STEP_OVERLAP): the bench's rotation of 8 worlds of 65 536
matches, a CUDA graph of 40 steps replayed until --steps launches have run, once with mode 3 (per-tile dependencies,
launches overlap) and once with mode 0 (grid-wide wait); the final states must agree bit for bit and the protocol's
error counter must read 0.  Also a single world chained with mode 2 against mode 0.

  python tools/overlap_soak.py [--steps 1000000]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rsoccer_b200 import _lib, engine as E  # noqa: E402


def run(mode, n_worlds, envs, steps, glen):
    dev = torch.device("cuda", 0)
    gen = torch.Generator().manual_seed(5)
    worlds, acts, outs = [], [], []
    for m in range(n_worlds):
        w = E.BatchedWorld(E.KIND_VSS, 0, 3, 3, 25, envs, device=dev, seed=77, env_offset=m * envs)
        w.set_option(_lib.OPT_STEP_OVERLAP, mode)
        w.task_reset(E.TASK_VSS_V0)
        worlds.append(w)
        acts.append((torch.rand(envs, 2, generator=gen) * 2 - 1).to(dev))
        outs.append(w.alloc_outputs(E.TASK_VSS_V0))
    st = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(st):
        for i in range(glen):                               # the first launches outside the graph
            worlds[i % n_worlds].vss_env_step(acts[i % n_worlds], out=outs[i % n_worlds])
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for i in range(glen):
                worlds[i % n_worlds].vss_env_step(acts[i % n_worlds], out=outs[i % n_worlds])
        reps = max(1, steps // glen)
        t0 = time.perf_counter()
        for _ in range(reps):
            g.replay()
        st.synchronize()
        dt = time.perf_counter() - t0
    errs = sum(w.get_option(_lib.OPT_OVERLAP_ERRORS) for w in worlds)
    states = [w.state.clone() for w in worlds]
    obs = [o[0].clone() for o in outs]
    for w in worlds:
        w.close()
    return states, obs, errs, reps * glen, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1000000)
    a = ap.parse_args()
    # small worlds: by default mode 3 picks the lane-per-match kernels and mode 0 the lane-per-body ones, two mappings
    # that agree to rounding, not to the bit -- the mapping is pinned so that only the protocol differs
    for what, n_worlds, envs, mode, glen, per_match in (
            ("8 worlds of 65 536 matches in rotation", 8, 65536, 3, 40, None),
            ("one world of 65 536 matches, chained", 1, 65536, 2, 40, None),
            ("125 worlds of 4 096 matches in rotation, lane per match", 125, 4096, 3, 500, "1"),
            ("125 worlds of 4 096 matches in rotation, lane per body", 125, 4096, 3, 500, "0")):
        if per_match is None:
            os.environ.pop("RS_PER_MATCH", None)
        else:
            os.environ["RS_PER_MATCH"] = per_match
        s1, o1, e1, n1, t1 = run(mode, n_worlds, envs, a.steps, glen)
        s0, o0, e0, n0, t0 = run(0, n_worlds, envs, a.steps, glen)
        same = all(torch.equal(x, y) for x, y in zip(s1, s0)) and all(torch.equal(x, y) for x, y in zip(o1, o0))
        print("SOAK %-58s mode %d: %d launches in %.1f s (%.2f us each), protocol errors %d; mode 0: %.1f s (%.2f us each); "
              "final states and observations bit-identical: %s" % (what, mode, n1, t1, t1 / n1 * 1e6, e1, t0, t0 / n0 * 1e6, same), flush=True)
        assert same and e1 == 0 and e0 == 0


if __name__ == "__main__":
    main()

"""e2e of the *_host steps with the actions read in place / staged with a copy (RS_OPT_HOST_COPY_ACTIONS).

usage: python tools/e2e_probe.py [--steps 60] [--configs vss65536,vss32768,vss4096,sd4096,cp16384]
For each config: us per host step for both settings, next to the plain D2H copy of the same block (the box's
ceiling), and a bit-for-bit check of the outputs and the raw state between the settings.
(profiles/r2_e2e_probe.txt also holds the run of the experimental chunk-pipelined host step, "plan" 2-5.)
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rsoccer_b200 import _lib, engine as E  # noqa: E402

CFG = {
    "vss": (E.KIND_VSS, 0, 3, 3, E.TASK_VSS_V0, 2),
    "sd": (E.KIND_SSL, 2, 1, 6, E.TASK_SSL_STATIC_DEFENDERS_V0, 5),
    "cp": (E.KIND_SSL, 2, 1, 1, E.TASK_SSL_CONTESTED_POSSESSION_V0, 5),
}


def make(name, n, seed=7):
    kind, ft, nb, ny, tid, ad = CFG[name]
    w = E.BatchedWorld(kind, ft, nb, ny, 25, n, device="cuda:0", seed=seed)
    w.task_reset(tid)
    return w, tid, ad


def host_step(w, name, tid, bufs):
    if name == "vss":
        w.vss_env_step_host(*bufs)
    else:
        w.ssl_env_step_host(tid, *bufs)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--configs", default="vss65536,vss32768,vss4096,sd4096,cp16384,sd65536")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    st = torch.cuda.Stream()
    for cfg in a.configs.split(","):
        name = cfg.rstrip("0123456789")
        n = int(cfg[len(name):])
        gen = torch.Generator().manual_seed(3)
        ref = None
        print("== %s" % cfg, flush=True)
        for copy_act in (0, 1):
            w, tid, ad = make(name, n)
            w.set_option(_lib.OPT_HOST_COPY_ACTIONS, copy_act)
            h_act = torch.empty(n, ad, dtype=torch.float32).pin_memory()
            gen.manual_seed(3)
            h_act.copy_(torch.rand(n, ad, generator=gen) * 2 - 1)
            out = w.alloc_host_outputs(tid)
            bufs = (h_act,) + tuple(out)
            with torch.cuda.stream(st):
                for _ in range(40):                      # same trajectory for every variant
                    host_step(w, name, tid, bufs)
                snap = [o.clone() for o in out] + [w.get_raw().cpu()]
                if ref is None:
                    ref = snap
                same = all(torch.equal(x.view(torch.uint8) if x.dtype != torch.uint8 else x,
                                       y.view(torch.uint8) if y.dtype != torch.uint8 else y) for x, y in zip(snap, ref))
                torch.cuda.synchronize()
                best = 1e9
                for _rep in range(3):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(st)
                    for _ in range(a.steps):
                        host_step(w, name, tid, bufs)
                    e1.record(st)
                    torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1) / a.steps * 1e3)
            print("  copy_actions %d: %8.2f us/step   %7.1f M env-steps/s   bits %s" %
                  (copy_act, best, n / best, "same" if same else "DIFFERENT"), flush=True)
            w.close()
        # ceiling: plain D2H of the same block
        od = {"vss": 40, "sd": 24, "cp": 14}[name]
        d2h = n * (od * 4 + 6)
        src = torch.empty(d2h, dtype=torch.uint8, device="cuda:0")
        dst = torch.empty(d2h, dtype=torch.uint8).pin_memory()
        with torch.cuda.stream(st):
            for _ in range(5):
                dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(a.steps):
                dst.copy_(src, non_blocking=True)
                st.synchronize()
            e1.record(st)
            torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / a.steps * 1e3
        print("  plain D2H of %d bytes + sync: %8.2f us  (%.1f GB/s)" % (d2h, us, d2h / us / 1e3), flush=True)


if __name__ == "__main__":
    main()

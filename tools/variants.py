#!/usr/bin/env python
"""Kernel-tuning harness: build one library per set of -D knobs, time them all in one GPU call.

  python tools/variants.py build  name=-DA,-DB ...     (here: nvcc cross-compiles into build/variants/)
  python tools/variants.py run [--task vss] [--sizes 4096,65536] [--mode 1]   (on the GPU box)

`run` times every library under build/variants/ with tools/step_timing.time_steps (one forked
child per library, RS_LIB selects it) and appends lines to gpurun_out/variants.txt.
The knobs are the RS_X_* (decomposition) and RS_O_* (candidate optimisation) macros of
rsoccer_b200/csrc/*.cuh; a product build defines none of them.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VDIR = os.path.join(ROOT, "build", "variants")
sys.path.insert(0, ROOT)


def build(specs):
    from rsoccer_b200 import _lib
    os.makedirs(VDIR, exist_ok=True)

    def one(spec):
        name, _, flags = spec.partition("=")
        out = os.path.join(VDIR, "lib_%s.so" % name)
        cmd = ["/usr/local/cuda/bin/nvcc"] + _lib.NVCC_FLAGS + [f for f in flags.split(",") if f] + ["-o", out, _lib.SOURCES[0]]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return name, r.returncode, r.stderr[-2000:]

    with ThreadPoolExecutor(max_workers=8) as ex:
        for name, rc, err in ex.map(one, specs):
            print(name, "ok" if rc == 0 else "FAILED\n" + err)


def run(argv):
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", default="vss")
    ap.add_argument("--sizes", default="4096,65536")
    ap.add_argument("--mode", default="1", help="RS_PER_MATCH value ('' = auto)")
    ap.add_argument("--only", default="")
    ap.add_argument("--steps", type=int, default=4000)
    ap.add_argument("--worlds", type=int, default=8, help="1 = state stays L2 resident")
    ap.add_argument("--env", default="", help="extra environment for the children, K=V,K=V")
    ap.add_argument("--graph-steps", type=int, default=8, help="steps per captured graph (step overlap drains at every replay)")
    a = ap.parse_args(argv)
    import torch  # noqa: F401  (imported before the fork: children pay only CUDA init)
    libs = sorted(f for f in os.listdir(VDIR) if f.endswith(".so"))
    if a.only:
        libs = [f for f in libs if any(k in f for k in a.only.split(","))]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "variants.txt"), "a")
    for f in libs:
        pid = os.fork()
        if pid == 0:
            os.environ["RS_LIB"] = os.path.join(VDIR, f)
            if a.mode != "":
                os.environ["RS_PER_MATCH"] = a.mode
            for kv in [x for x in a.env.split(",") if x]:
                os.environ[kv.split("=")[0]] = kv.split("=")[1]
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import step_timing as T
            res = []
            for n in [int(x) for x in a.sizes.split(",")]:
                try:
                    res.append("%d:%.2f" % (n, T.time_steps(a.task, n, n_worlds=a.worlds, steps=a.steps, graph_steps=a.graph_steps)))
                except Exception as ex:  # noqa: BLE001
                    res.append("%d:ERR(%s)" % (n, str(ex)[:80]))
            line = "%-28s task=%s mode=%s worlds=%d %s  %s" % (f[4:-3], a.task, a.mode, a.worlds, a.env, "  ".join(res))
            print(line, flush=True)
            log.write(line + "\n")
            log.flush()
            os._exit(0)
        os.waitpid(pid, 0)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    else:
        run(sys.argv[2:])

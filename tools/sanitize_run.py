#!/usr/bin/env python
"""Small driver for compute-sanitizer (memcheck / racecheck / initcheck): every fused task step,
both kernel mappings, ragged sizes, short episodes so that the auto-reset paths run, crowded scenes so that
the shared-memory contact resolve runs, and every step-overlap mode with back-to-back launches on a fixed
action buffer so that the tile hand-over runs; the host-buffer steps (blocking and split-phase, actions in place /
staged / pageable).

  compute-sanitizer --tool racecheck python tools/sanitize_run.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity import random_raw  # noqa: E402
from rsoccer_b200 import _lib, engine as E  # noqa: E402

TASKS = ((0, 0, 0, 3, 3, 2), (1, 1, 2, 1, 6, 5), (2, 1, 2, 1, 1, 5), (3, 1, 2, 1, 4, 4), (4, 1, 2, 2, 0, 3))
rng = np.random.default_rng(0)
for mode, packed, overlap in (("1", "1", 0), ("1", "0", 0), ("0", "0", 0), ("1", "1", 2), ("1", "1", 3), ("0", "0", 2), ("1", "1", 1)):
    os.environ["RS_PER_MATCH"] = mode
    os.environ["RS_PACKED"] = packed         # VSS-v0 lane-per-match kernel: packed fp32x2 forms + pair table / scalar forms
    for task, kind, ft, nb, ny, nact in TASKS:
        if packed == "1" and task != 0 and overlap == 0:
            continue
        for n in (200, 1):
            w = E.BatchedWorld(kind, ft, nb, ny, 25, n, seed=3)
            w.set_option(_lib.OPT_STEP_OVERLAP, overlap)
            w.task_reset(task)
            g = torch.Generator().manual_seed(task)
            fp = w.field_params()
            out = w.alloc_outputs(task)
            for it in range(8):
                a = (torch.rand(n, nact, generator=g) * 2 - 1).cuda()
                if it == 4:     # a crowded scene: contacts in most matches
                    w.set_raw(random_raw(rng, n, nb + ny, fp["length"] / 2 - 0.1, fp["width"] / 2 - 0.1, crowd=0.8).astype(np.float32))
                for rep in range(3 if overlap else 1):      # back to back on one action buffer: chained launches
                    if task == 0:
                        w.vss_env_step(a, max_steps=3, out=out)
                    else:
                        w.ssl_env_step(task, a, max_steps=3, out=out)
            assert w.get_option(_lib.OPT_OVERLAP_ERRORS) == 0
            # host-buffer steps: blocking (pinned actions in place / staged, pageable), then split-phase
            h_out = w.alloc_host_outputs(task)
            for pinned, copy_act in ((True, -1), (True, 1 if task else 0), (False, -1)):
                w.set_option(_lib.OPT_HOST_COPY_ACTIONS, copy_act)
                h_act = torch.rand(n, nact, generator=g) * 2 - 1
                if pinned:
                    h_act = h_act.pin_memory()
                if task == 0:
                    w.vss_env_step_host(h_act, *h_out, max_steps=3)
                    w.vss_env_step_host_begin(h_act, *h_out, max_steps=3)
                else:
                    w.ssl_env_step_host(task, h_act, *h_out, max_steps=3)
                    w.ssl_env_step_host_begin(task, h_act, *h_out, max_steps=3)
                w.host_step_wait()
            c = torch.rand(n, nb + ny, w.cmd_dim, generator=g).cuda()
            w.step(c)
            w.get_state()
            torch.cuda.synchronize()
            w.close()
print("sanitize_run: done")

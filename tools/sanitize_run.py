#!/usr/bin/env python
"""Small driver for compute-sanitizer (memcheck / racecheck / initcheck): every fused task step,
both kernel mappings, ragged sizes, short episodes so that the auto-reset paths run.

  compute-sanitizer --tool racecheck python tools/sanitize_run.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from rsoccer_b200 import engine as E  # noqa: E402

TASKS = ((0, 0, 0, 3, 3, 2), (1, 1, 2, 1, 6, 5), (2, 1, 2, 1, 1, 5), (3, 1, 2, 1, 4, 4), (4, 1, 2, 2, 0, 3))
for mode, packed in (("1", "1"), ("1", "0"), ("0", "0")):
    os.environ["RS_PER_MATCH"] = mode
    os.environ["RS_PACKED"] = packed         # VSS-v0 lane-per-match kernel: packed fp32x2 forms + pair table / scalar forms
    for task, kind, ft, nb, ny, nact in TASKS:
        if packed == "1" and task != 0:
            continue
        for n in (200, 1):
            w = E.BatchedWorld(kind, ft, nb, ny, 25, n, seed=3)
            w.task_reset(task)
            g = torch.Generator().manual_seed(task)
            for _ in range(8):
                a = (torch.rand(n, nact, generator=g) * 2 - 1).cuda()
                if task == 0:
                    w.vss_env_step(a, max_steps=3)
                else:
                    w.ssl_env_step(task, a, max_steps=3)
            c = torch.rand(n, nb + ny, w.cmd_dim, generator=g).cuda()
            w.step(c)
            w.get_state()
            torch.cuda.synchronize()
            w.close()
print("sanitize_run: done")

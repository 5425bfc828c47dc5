#!/usr/bin/env python
"""Contact statistics of VSS-v0 under the reference's OU-driven robots (CPU oracle).

Used to size the divergence of the pair / wall phases: per lane a
contact is rare, per warp of 32 matches it is the common case."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O  # noqa: E402


def main(n=2048, steps=2400):
    o = O.OracleWorld(0, 0, 3, 3, 25, n, seed=1, threads=O.max_threads())
    o.task_reset(O.TASK_VSS)
    rng = np.random.default_rng(0)
    rad = np.array([0.0215] + [0.0375] * 6)
    for t in range(steps + 1):
        if t in (0, 50, 200, 600, 1199, 1200, 2400):
            raw = o.get_raw()
            pos = np.zeros((n, 7, 2))
            pos[:, 0] = raw[:, 0:2]
            for r in range(6):
                pos[:, r + 1] = raw[:, 4 + 6 * r:6 + 6 * r]
            cnt = np.zeros(n, int)
            for i in range(7):
                for j in range(i + 1, 7):
                    cnt += np.linalg.norm(pos[:, i] - pos[:, j], axis=1) < rad[i] + rad[j] + 1e-4
            wall = (np.abs(pos[:, :, 0]) > 0.75 - rad - 1e-4) | (np.abs(pos[:, :, 1]) > 0.65 - rad - 1e-4)
            p_any = (cnt > 0).mean()
            print("t=%4d  pair contacts/env %.3f  P(env has one) %.3f  P(warp of 32 has one) %.3f  "
                  "robots at a wall %.3f" % (t, cnt.mean(), p_any, 1 - (1 - p_any) ** 32, wall[:, 1:].mean()))
        o.vss_env_step(rng.uniform(-1, 1, (n, 2)).astype(np.float32))


if __name__ == "__main__":
    main()

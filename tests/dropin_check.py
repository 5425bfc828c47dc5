#!/usr/bin/env python
"""Drop-in check: the UNMODIFIED reference env classes (gym.make of the five registered ids, the
README.md:116-133 loop) running on the CUDA engine through the `robosim`-compatible module
(rsoccer_b200.compat.robosim; reference seam: rsoccer_gym/Simulators/rsim.py:2, 116-124, 169-177),
next to the same classes running on the CPU oracle.  Before every step the CUDA-backed env is
re-synced to the oracle-backed one (fp32-rounded raw state, OU state, episode counters -- the
scheme of tests/golden/make_golden.py), both get the same action and the same np.random stream, and
observation / reward / done must agree to the documented tolerance (1e-4, velocities scaled by
max(1, L/2), tests/parity.py) unless the oracle flags the step as decided by < 5e-6 m.

    python tests/dropin_check.py [steps]      needs a GPU and the reference package
                                              (baseline/_ref, installed by __graft_entry__.build(), or /root/reference)
"""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "golden", "shims"))


def reference_path():
    for p in (os.environ.get("RS_REFERENCE"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if p and os.path.isdir(os.path.join(p, "rsoccer_gym")):
            return p
    return None


def main(steps=80):
    ref = reference_path()
    if ref is None:
        print("SKIP: no reference package (baseline/_ref)")
        return 0
    import robosim as oracle_robosim                      # tests/golden/shims: served by the oracle
    from rsoccer_b200 import compat
    selfcheck = os.environ.get("RS_DROPIN_SELFCHECK") == "1"      # CPU: the harness against itself (oracle twice)
    if selfcheck:
        cuda_robosim = oracle_robosim
    else:
        from rsoccer_b200.compat import robosim as cuda_robosim
    compat.install(robosim_module=cuda_robosim)           # gymnasium / pygame stand-ins only when absent
    sys.path.insert(0, ref)
    import gymnasium as gym
    import rsoccer_gym  # noqa: F401
    import rsoccer_gym.Simulators.rsim as rsim_mod
    if not selfcheck:
        import torch

    def make(eid, backend):
        rsim_mod.robosim = backend                        # looked up at call time by _init_simulator
        return gym.make(eid)

    SYNC = ("steps", "previous_ball_potential", "checkpoints_count", "stopped_steps", "holding_steps")
    report = []
    for eid in ("VSS-v0", "SSLStaticDefenders-v0", "SSLContestedPossession-v0", "SSLDribbling-v0",
                "SSLPassEndurance-v0"):
        random.seed(7); np.random.seed(7)
        rng = np.random.default_rng(7)
        eo, eg = make(eid, oracle_robosim), make(eid, cuda_robosim)
        uo, ug = eo.unwrapped, eg.unwrapped
        assert selfcheck or type(ug.rsim.simulator).__module__.endswith("compat.robosim"), "CUDA engine not behind the env"
        fo, fg = uo.rsim.get_field_params(), ug.rsim.get_field_params()
        assert fo == fg, (eid, fo, fg)
        eo.reset(); eg.reset()
        vel_scale = max(1.0, fo.length / 2)
        n_act = eo.action_space.shape[0]
        worst, flagged, dones = 0.0, 0, 0
        for t in range(steps):
            # ---- re-sync the CUDA-backed env to the oracle-backed one
            raw = uo.rsim.simulator._w.get_raw()[0].astype(np.float32)
            uo.rsim.simulator._w.set_raw(raw.astype(np.float64).reshape(1, -1))
            ug.rsim.simulator._w.set_raw(raw.astype(np.float64).reshape(1, -1) if selfcheck else torch.from_numpy(raw).reshape(1, -1))
            uo.frame, ug.frame = uo.rsim.get_frame(), ug.rsim.get_frame()
            for k in SYNC:
                if hasattr(uo, k):
                    setattr(ug, k, getattr(uo, k))
            if hasattr(uo, "ou_actions"):
                for a, b in zip(uo.ou_actions, ug.ou_actions):
                    b.x_prev = np.array(a.x_prev, copy=True)
            act = rng.uniform(-1, 1, n_act).astype(np.float32)
            if eid != "VSS-v0" and t % 3 == 0:            # chase the ball so that contacts / kicks / dribbling happen
                d = raw[0:2] - raw[4:6]
                act[0:2] = 0.6 * d / (np.linalg.norm(d) + 1e-9)
            st = np.random.get_state()
            oo, ro, do, tro, _ = eo.step(act.copy())
            np.random.set_state(st)
            og, rg, dg, trg, _ = eg.step(act.copy())
            ok = uo.rsim.simulator._w.margin()[0] >= 5e-6
            if not ok:
                flagged += 1
            else:
                tol = np.full(oo.shape, 1e-4 * vel_scale)
                err = np.abs(np.asarray(oo, dtype=np.float64) - np.asarray(og, dtype=np.float64))
                if eid != "VSS-v0":
                    # the reference divides v_theta in deg/s by 10: 1e-4 rad/s is 5.7e-4 there
                    tol = np.maximum(tol, 1e-4 * vel_scale * 57.3 / 10.0 * (err > 0))
                assert (err <= tol).all(), (eid, t, float(err.max()), np.argmax(err - tol))
                assert abs(float(ro) - float(rg)) <= 1e-4 * vel_scale * 10, (eid, t, ro, rg)
                assert bool(do) == bool(dg) and bool(tro) == bool(trg), (eid, t, do, dg)
                worst = max(worst, float(err.max()))
            if do or tro:
                dones += 1
                eo.reset(); eg.reset()
        assert flagged <= steps // 4, (eid, flagged)      # dribbling starts with the ball resting exactly on the mouth
        report.append("%s: %d steps, max |obs_cuda - obs_oracle| = %.2e, %d episode ends, %d flagged"
                      % (eid, steps, worst, dones, flagged))
        eo.close(); eg.close()
    print("\n".join(report))
    print("OK")
    return 0


if __name__ == "__main__":
    sys.exit(main(int(sys.argv[1]) if len(sys.argv) > 1 else 80))

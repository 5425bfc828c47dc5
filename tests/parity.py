"""Shared helpers of the parity tests: seeded scenes, oracle <-> CUDA comparison.

The CUDA path is fp32, the oracle fp64.  One control step from IDENTICAL (fp32-rounded)
inputs must agree to 1e-4 abs on every state component (north_star tolerance), except in
envs the oracle itself flags as ill-conditioned: a discrete decision (contact / no contact,
wall hit, kicker box, goal line, deadzone) whose margin is below MARGIN_EPS, where a
1-ulp input difference legitimately flips the branch.  Those are counted and bounded.
"""
import numpy as np

TOL = 1e-4
MARGIN_EPS = 5e-6


def wrap(a):
    return (a + np.pi) % (2 * np.pi) - np.pi


def random_raw(rng, n, R, half_len, half_wid, v_ball=1.5, v_rbt=1.0, w_rbt=10.0, crowd=0.3):
    """[n, 4 + 6R] raw states: bodies uniformly in the field, a fraction `crowd` of the envs
    squeezed into a small box so that contacts are frequent."""
    raw = np.zeros((n, 4 + 6 * R))
    scale = np.where(rng.random(n) < crowd, 0.4, 1.0)[:, None]
    half_len = half_len - 0.03   # keep centres out of the solid corner / goal-wall boxes
    raw[:, 0] = rng.uniform(-half_len, half_len, n) * scale[:, 0]
    raw[:, 1] = rng.uniform(-half_wid, half_wid, n) * scale[:, 0]
    raw[:, 2:4] = rng.normal(0, v_ball, (n, 2))
    for r in range(R):
        o = 4 + 6 * r
        raw[:, o] = rng.uniform(-half_len, half_len, n) * scale[:, 0]
        raw[:, o + 1] = rng.uniform(-half_wid, half_wid, n) * scale[:, 0]
        raw[:, o + 2] = rng.uniform(-np.pi, np.pi, n)
        raw[:, o + 3:o + 5] = rng.normal(0, v_rbt, (n, 2))
        raw[:, o + 5] = rng.normal(0, w_rbt, n)
    return raw.astype(np.float32).astype(np.float64)


def random_cmds(rng, kind, n, R):
    """robosim.step command rows: VSS wheel rad/s; SSL local twists, a fifth as raw wheel speeds, kicks, dribblers"""
    if kind == 0:
        c = rng.uniform(-60, 60, (n, R, 2))
        c[rng.random((n, R)) < 0.2] = 0.0
        return c.astype(np.float32)
    c = np.zeros((n, R, 8), dtype=np.float32)
    c[:, :, 1:3] = rng.uniform(-2.5, 2.5, (n, R, 2))
    c[:, :, 3] = rng.uniform(-10, 10, (n, R))
    ws = rng.random((n, R)) < 0.2
    c[ws, 0] = 1.0
    c[ws, 1:5] = rng.uniform(-150, 150, (int(ws.sum()), 4))
    c[:, :, 5] = np.where(rng.random((n, R)) < 0.3, 5.0, 0.0)
    c[:, :, 7] = (rng.random((n, R)) < 0.4).astype(np.float32)
    c[rng.random((n, R)) < 0.2] = 0.0
    return c


# (kind, field_type, n_blue, n_yellow) of test_step_parity_resynced -> the largest fraction of its 6 x n
# contact-rich random scenes the oracle flags as ill conditioned at eps = 2e-5 (measured on the CPU, it does
# not depend on the GPU: tests/test_oracle.py::test_flagged_fraction_of_the_resynced_scenes) + 2 points.
# What is left are injected overlaps no trajectory reaches (a ball inside a robot, a centre inside a goal wall).
RESYNC_CASES = [(0, 0, 3, 3), (0, 1, 5, 5), (0, 0, 1, 1), (0, 0, 2, 3), (1, 2, 1, 6), (1, 2, 1, 1), (1, 0, 3, 3),
                (1, 2, 1, 0), (1, 2, 1, 4), (1, 2, 2, 0), (1, 1, 11, 11)]
RESYNC_EPS = 2e-5
RESYNC_MAX_FLAGGED = {(0, 0, 3, 3): 0.056, (0, 1, 5, 5): 0.068, (0, 0, 1, 1): 0.032, (0, 0, 2, 3): 0.051,
                      (1, 2, 1, 6): 0.067, (1, 2, 1, 1): 0.045, (1, 0, 3, 3): 0.051, (1, 2, 1, 0): 0.046,
                      (1, 2, 1, 4): 0.061, (1, 2, 2, 0): 0.045, (1, 1, 11, 11): 0.070}


def resynced_scene(rng, kind, n, R, fp, it):
    """scene `it` of test_step_parity_resynced: (raw [n, 4 + 6R] fp32-exact, cmds)"""
    raw = random_raw(rng, n, R, fp["length"] / 2 + 0.05, fp["width"] / 2,
                     v_ball=2.0 if kind else 1.0, v_rbt=1.0, w_rbt=6.0)
    if kind == 1 and it % 2 == 1:   # put the ball in front of robot 0's mouth in half the envs
        k = rng.random(n) < 0.5
        th = raw[:, 4 + 2]
        d = rng.uniform(0.085, 0.115, n)
        lat = rng.uniform(-0.05, 0.05, n)
        raw[k, 0] = (raw[:, 4] + np.cos(th) * d - np.sin(th) * lat)[k]
        raw[k, 1] = (raw[:, 5] + np.sin(th) * d + np.cos(th) * lat)[k]
    if kind == 0:
        # VSS: beyond the goal line only the goal mouth is open; a centre next to the solid corner is pulled
        # back to 3 cm from its face (a robot then overlaps the wall by 7.5 mm: wall contacts, not centres
        # inside the solid, which no trajectory reaches and the oracle can only call ill conditioned)
        gw, hl = fp["goal_width"] / 2, fp["length"] / 2
        for b in range(R + 1):
            cx, cy = (0, 1) if b == 0 else (4 + 6 * (b - 1), 5 + 6 * (b - 1))
            solid = np.abs(raw[:, cy]) > gw - 0.045
            raw[solid, cx] = np.clip(raw[solid, cx], -(hl - 0.03), hl - 0.03)
    raw = raw.astype(np.float32).astype(np.float64)
    return raw, random_cmds(rng, kind, n, R)


def raw_diff(a, b, R, vel_scale=1.0):
    """max abs difference per env between two raw states (angles modulo 2 pi).

    vel_scale divides the velocity errors: 1 for VSS (the 1e-4 north_star bound as is).  On
    the 6-12 m SSL fields a contact normal is a difference of O(L/2) fp32 coordinates, so
    the post-contact velocity carries ulp(L/2)/ulp(1 m) more rounding; SSL tests pass
    vel_scale = max(1, L/2 [m]).  Positions and angles are never scaled."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = np.abs(a - b)
    d[:, 2:4] /= vel_scale
    for r in range(R):
        o = 4 + 6 * r
        d[:, o + 2] = np.abs(wrap(a[:, o + 2] - b[:, o + 2]))
        d[:, o + 3:o + 6] /= vel_scale
    return d.max(axis=1)


def check_close(err, margin, what, tol=TOL, max_flagged=0.05, eps=MARGIN_EPS):
    """err, margin: per-env arrays.  Every env with margin >= eps must be within tol."""
    err = np.asarray(err)
    ok = margin >= eps
    frac_flagged = 1.0 - ok.mean()
    assert frac_flagged <= max_flagged, "%s: %.3f%% of envs ill-conditioned" % (what, 100 * frac_flagged)
    worst = err[ok].max() if ok.any() else 0.0
    assert worst <= tol, "%s: max abs err %.3e > %.1e (env %d, margin %.2e)" % (
        what, worst, tol, int(np.argmax(np.where(ok, err, -1))), margin[int(np.argmax(np.where(ok, err, -1)))])
    return worst, frac_flagged

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

"""Opportunistic comparison with the REAL engine: runs only where `robosim` (rc-robosim >= 1.2.0, the
pybind11 module the reference imports at rsoccer_gym/Simulators/rsim.py:2, pinned in setup.py:15) can be
imported.  It cannot in this image, so here every test below is skipped and the
physics stays "parity unpinned"; the day a wheel is installed these tests pin (a) every [MEM] field
constant of include/rs_spec.h against get_field_params() (rsim.py:49-50, Entities/Field.py:4-21) and
(b) report the per-step divergence of the 2-D model from rSim on identical commands."""
import numpy as np
import pytest

robosim = pytest.importorskip("robosim", reason="rc-robosim is not installed (not installable in this image)")

WORLDS = [("VSS", 0, 0, 3, 3), ("VSS", 0, 1, 5, 5), ("SSL", 1, 0, 6, 6), ("SSL", 1, 1, 11, 11), ("SSL", 1, 2, 1, 6)]


def _dummy(nb, ny):
    # rsim.py:19-24
    return [0.0, 0.0, 0.0, 0.0], [[-0.2 * i, 0.0, 0.0] for i in range(1, nb + 1)], [[0.2 * i, 0.0, 0.0] for i in range(1, ny + 1)]


@pytest.mark.parametrize("cls,kind,ft,nb,ny", WORLDS)
def test_field_params_equal_rsim(oracle, cls, kind, ft, nb, ny):
    """the 17 Field values: every [MEM] constant (wheel radius, rpm, kicker geometry, wheel angles, the
    field_type -> dimensions table) is confirmed or refuted here"""
    sim = getattr(robosim, cls)(ft, nb, ny, 25, *_dummy(nb, ny))
    ref = dict(sim.get_field_params())
    ours = oracle.OracleWorld(kind, ft, nb, ny, 25, 1).field_params()
    assert set(ref) == set(ours)
    bad = {k: (ref[k], ours[k]) for k in ref if abs(ref[k] - ours[k]) > 1e-9}
    assert not bad, bad


@pytest.mark.parametrize("cls,kind,ft,nb,ny", WORLDS[:1] + WORLDS[-1:])
@pytest.mark.xfail(strict=False, reason="parity unpinned: a 2-D impulse model against rSim's 3-D ODE world; the numbers are the point")
def test_100_step_rollout_against_rsim(oracle, cls, kind, ft, nb, ny):
    """same reset, same commands, 100 control steps, state re-synced from rSim before every step: per-step
    |x, y, theta, v| difference against the north_star's 1e-4"""
    R = nb + ny
    rng = np.random.default_rng(0)
    sim = getattr(robosim, cls)(ft, nb, ny, 25, *_dummy(nb, ny))
    o = oracle.OracleWorld(kind, ft, nb, ny, 25, 1)
    K = 6 if kind == 0 else 11
    worst = 0.0
    for t in range(100):
        st = np.asarray(sim.get_state(), dtype=np.float64)
        raw = np.zeros(4 + 6 * R)
        raw[0:2], raw[2:4] = st[0:2], st[3:5]
        for r in range(R):
            q = st[5 + K * r:5 + K * r + 6]
            raw[4 + 6 * r:10 + 6 * r] = (q[0], q[1], np.deg2rad(q[2]), q[3], q[4], np.deg2rad(q[5]))
        o.set_raw(raw.reshape(1, -1))
        if kind == 0:
            cmd = rng.uniform(-30, 30, (R, 2))
        else:
            cmd = np.zeros((R, 8)); cmd[:, 1:3] = rng.uniform(-1.5, 1.5, (R, 2)); cmd[:, 3] = rng.uniform(-5, 5, R)
        sim.step(cmd); o.step(cmd.reshape(1, R, -1))
        a, b = np.asarray(sim.get_state(), dtype=np.float64), o.get_state()[0]
        cols = [0, 1, 3, 4] + [5 + K * r + c for r in range(R) for c in (0, 1, 3, 4)]
        worst = max(worst, float(np.abs(a[cols] - b[cols]).max()))
    print("%s %dv%d: max per-step |x, y, v| difference to rSim over 100 re-synced steps = %.3e" % (cls, nb, ny, worst))
    assert worst < 1e-4

"""include/rs_spec.h (rs_params_fill) is shared by the CUDA library and by the oracle, so no CUDA <-> oracle test
can see a wrong derived constant in it.  This file re-derives the block WITHOUT the header: numpy for the omni
wheel matrices (np.linalg.pinv, not the header's cofactor inverse) and plain arithmetic / hand-computed
literals for the walls, the kicker box, the wheel speed limit and the env normalisers, for every field type.

Each input constant carries its provenance: [REF file:line] = stated by the reference tree, [MEM] = the
builder's recollection of rSim's config (what a real `get_field_params()` would confirm: Entities/Field.py:4-21;
tests/test_robosim_optional.py does that comparison the day robosim imports), [SPEC] = a choice of this model."""
import numpy as np
import pytest

# (kind, field_type) -> length, width, penalty_length, penalty_width, goal_width, goal_depth
FIELDS = {
    (0, 0): (1.5, 1.3, 0.15, 0.7, 0.4, 0.1),        # [REF Render/field.py:189-199] VSS 3 v 3
    (0, 1): (2.2, 1.8, 0.15, 0.8, 0.4, 0.15),       # [MEM] VSS 5 v 5 (vss/README.md:4 names the type only)
    (1, 0): (9.0, 6.0, 1.0, 2.0, 1.0, 0.18),        # [REF Render/field.py:252-263] SSL div B
    (1, 1): (12.0, 9.0, 1.8, 3.6, 1.8, 0.18),       # [MEM] SSL div A
    (1, 2): (6.0, 4.0, 0.8, 1.8, 0.8, 0.18),        # [MEM] 2021 hardware-challenge field (ssl/README.md:4)
}
ROBOT = {
    # ball_radius, rbt_radius, wheel_radius, max_rpm, kicker distance / thickness / width, wheel angles [deg]
    0: (0.0215,     # [REF Render/ball.py:6]
        0.0375,     # [REF vss_gym_base.py:57 comment]
        0.026,      # [MEM]
        440.0,      # [MEM]
        0.0, 0.0, 0.0, (90.0, 270.0, 0.0, 0.0)),
    1: (0.0215,     # [MEM]
        0.09,       # [REF ssl_gym_base.py:58 comment]
        0.02475,    # [MEM]
        160.0 * 60.0 / (2.0 * np.pi),   # 160 rad/s [REF static_defenders.py:71]
        0.073, 0.005, 0.08,             # [MEM]
        (60.0, 135.0, 225.0, 300.0)),   # [MEM]
}


@pytest.mark.parametrize("kind,ft", sorted(FIELDS))
def test_derived_block_matches_an_independent_derivation(oracle, kind, ft):
    nb, ny = (3, 3) if kind == 0 else (1, 6)
    w = oracle.OracleWorld(kind, ft, nb, ny, 25, 1)
    P, F = w.params(), w.field_params()
    L, W, pl, pw, gw, gd = FIELDS[(kind, ft)]
    br, rr, rw, rpm, dk, kt, kw, ang = ROBOT[kind]
    assert [F[k] for k in ("length", "width", "penalty_length", "penalty_width", "goal_width", "goal_depth")] == [L, W, pl, pw, gw, gd]
    assert (F["ball_radius"], F["rbt_radius"], F["rbt_wheel_radius"]) == (br, rr, rw)
    assert abs(F["rbt_motor_max_rpm"] - rpm) < 1e-12
    assert (F["rbt_distance_center_kicker"], F["rbt_kicker_thickness"], F["rbt_kicker_width"]) == (dk, kt, kw)
    assert tuple(F["rbt_wheel%d_angle" % i] for i in range(4)) == ang
    assert P["dt"] == 0.025 and abs(P["h"] - 0.005) < 1e-18
    assert abs(P["wheel_max_rad_s"] - rpm * 2 * np.pi / 60) < 1e-12
    assert abs(P["ball_decel"] - 0.05 * 9.81) < 1e-15                    # [MEM] mu_roll 0.05, g 9.81
    if kind == 0:
        # walls: field rectangle + goal recess = [0, L/2 + gd] x [0, W/2] minus the corner [L/2, inf) x [gw/2, inf)
        assert (P["x_out"], P["y_out"], P["x_near"], P["n_box"]) == (L / 2 + gd, W / 2, L / 2, 1)
        assert tuple(P["box"][0][:2]) == (L / 2, gw / 2) and P["box"][0][2] >= 1e8 and P["box"][0][3] >= 1e8
        assert P["half_track"] == 0.04                                   # [REF vss_gym_base.py:57-58]
        assert (P["ball_mass"], P["rbt_mass"]) == (0.046, 0.18)          # [MEM]
    else:
        m, t = 0.7, 0.02                                                 # [MEM] field + referee margin, goal wall thickness
        assert abs(P["x_out"] - (L / 2 + m)) < 1e-12 and abs(P["y_out"] - (W / 2 + m)) < 1e-12 and P["n_box"] == 2
        assert np.allclose(P["box"][0], [L / 2, gw / 2, L / 2 + gd + t, gw / 2 + t], atol=1e-12)          # goal side wall
        assert np.allclose(P["box"][1], [L / 2 + gd, -(gw / 2 + t), L / 2 + gd + t, gw / 2 + t], atol=1e-12)  # back wall
        assert (P["ball_mass"], P["rbt_mass"]) == (0.043, 2.2)           # [MEM]
        # omni kinematics: wheel i at angle a_i drives along (-sin a_i, cos a_i) at lever arm rr
        a = np.deg2rad(np.array(ang))
        J = np.stack([-np.sin(a), np.cos(a), np.full(4, rr)], axis=1)
        assert np.abs(P["omni_J"] - J).max() < 1e-14
        assert np.abs(P["omni_Jpinv"] - np.linalg.pinv(J)).max() < 1e-12
        assert np.abs(P["omni_Jpinv"] @ J - np.eye(3)).max() < 1e-12
        # grSim-style touching box in front of the flat mouth, the mouth chord
        assert abs(P["kick_centre"] - (dk + kt / 2)) < 1e-15 and abs(P["kick_reach"] - (2 * kt + br)) < 1e-15
        assert abs(P["kick_half_width"] - kw / 2) < 1e-15
        assert abs(P["mouth_half_chord"] - np.sqrt(rr * rr - dk * dk)) < 1e-12
        assert P["kick_speed_max"] == 6.5                                # [SPEC]


def test_hand_computed_literals():
    """numbers worked out by hand (no formula shared with the header), for the three benchmarked worlds"""
    from oracle import oracle as O
    v = O.OracleWorld(0, 0, 3, 3, 25, 1).params()
    assert abs(v["wheel_max_rad_s"] - 46.07669225) < 1e-7                # 440 rpm = 440 * 0.10471976 rad/s
    assert (v["x_out"], v["y_out"]) == (0.85, 0.65) and tuple(v["box"][0][:2]) == (0.75, 0.2)
    s = O.OracleWorld(1, 2, 1, 6, 25, 1).params()
    assert np.allclose(s["omni_J"][0], [-0.8660254038, 0.5, 0.09], atol=1e-9)          # wheel 0 at 60 deg
    assert np.allclose(s["omni_J"][1], [-0.7071067812, -0.7071067812, 0.09], atol=1e-9)  # wheel 1 at 135 deg
    # J+ row of the forward speed: by symmetry (+-a, +-b) with a = sin60 / (2 (sin^2 60 + sin^2 45)) etc.
    assert np.allclose(s["omni_Jpinv"][0], [-0.3464101615, -0.2828427125, 0.2828427125, 0.3464101615], atol=1e-9)
    assert abs(s["mouth_half_chord"] - 0.0526402888) < 1e-9              # sqrt(0.0081 - 0.005329)
    assert (round(s["kick_centre"], 6), round(s["kick_reach"], 6), s["kick_half_width"]) == (0.0755, 0.0315, 0.04)
    assert abs(s["x_out"] - 3.7) < 1e-12 and abs(s["y_out"] - 2.7) < 1e-12
    assert np.allclose(s["box"], [[3.0, 0.4, 3.2, 0.42], [3.18, -0.42, 3.2, 0.42]], atol=1e-12)


def test_env_normalisers_follow_the_reference_formulas(oracle):
    """vss_gym_base.py:52-58 / ssl_gym_base.py:53-59 evaluated here on the Field values: max_pos, max_v, max_w --
    and against the first observation the oracle's task layer builds from a known state"""
    w = oracle.OracleWorld(0, 0, 3, 3, 25, 1)
    F = w.field_params()
    max_pos = max(F["width"] / 2, F["length"] / 2 + F["penalty_length"])
    max_v = F["rbt_motor_max_rpm"] / 60 * 2 * np.pi * F["rbt_wheel_radius"]
    max_w = np.rad2deg(max_v / 0.04)
    assert abs(max_pos - 0.9) < 1e-15 and abs(max_v - 1.1979939985) < 1e-9 and abs(max_w - 1716.0) < 0.05
    raw = np.zeros((1, 4 + 36))
    raw[0, :4] = (0.45, -0.09, 0.5990, -0.2396)
    raw[0, 4:10] = (0.18, 0.27, np.pi / 6, 0.11979939985, 0.0, np.deg2rad(171.6))
    w.set_raw(raw)
    obs = w.task_obs(oracle.TASK_VSS)[0]
    assert np.allclose(obs[:4], [0.5, -0.1, 0.5990 / max_v, -0.2396 / max_v], atol=1e-9)
    assert np.allclose(obs[4:11], [0.2, 0.3, 0.5, np.sqrt(3) / 2, 0.1, 0.0, 171.6 / max_w], atol=1e-9)

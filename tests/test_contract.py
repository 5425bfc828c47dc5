"""Behavioural contract K1-K12: what the reference envs rely on from
`robosim`, each item derived from reference usage (file:line in the docstrings).  Run on the
CPU oracle always and on the CUDA engine under -m gpu (through the same C ABI the product uses).
"""
import numpy as np
import pytest


class Sim:
    """uniform numpy facade over OracleWorld / BatchedWorld"""

    def __init__(self, backend, mod, kind, ft, nb, ny, n=1):
        self.backend, self.nb, self.ny, self.R, self.n = backend, nb, ny, nb + ny, n
        self.kind = kind
        if backend == "oracle":
            self.w = mod.OracleWorld(kind, ft, nb, ny, 25, n)
        else:
            self.w = mod.BatchedWorld(kind, ft, nb, ny, 25, n)
        self.K = 6 if kind == 0 else 11
        self.C = 2 if kind == 0 else 8

    def field(self):
        return self.w.field_params()

    def reset(self, ball, blue, yellow):
        f = np.float64 if self.backend == "oracle" else np.float32
        self.w.reset(np.asarray(ball, f).reshape(self.n, 4), np.asarray(blue, f).reshape(self.n, self.nb, 3),
                     np.asarray(yellow, f).reshape(self.n, self.ny, 3) if self.ny else np.zeros((self.n, 0, 3), f))

    def step(self, cmds, k=1):
        f = np.float64 if self.backend == "oracle" else np.float32
        c = np.asarray(cmds, f).reshape(self.n, self.R, self.C)
        for _ in range(k):
            self.w.step(c)

    def state(self):
        s = self.w.get_state()
        return np.asarray(s.cpu().numpy() if hasattr(s, "cpu") else s, dtype=np.float64)

    def ball(self):
        return self.state()[:, :5]

    def robot(self, r):
        return self.state()[:, 5 + self.K * r:5 + self.K * (r + 1)]


@pytest.fixture(params=["oracle", pytest.param("cuda-per_match", marks=pytest.mark.gpu),
                        pytest.param("cuda-per_body", marks=pytest.mark.gpu)])
def mk(request, monkeypatch):
    if request.param == "oracle":
        from oracle import oracle as O
        O.build()
        return lambda *a, **k: Sim("oracle", O, *a, **k)
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from rsoccer_b200 import engine as E
    # both kernel mappings (rs_create reads RS_PER_MATCH; unset = chosen by world size)
    monkeypatch.setenv("RS_PER_MATCH", "1" if request.param.endswith("per_match") else "0")
    return lambda *a, **k: Sim("cuda", E, *a, **k)


VSS_BLUE = [[-0.4, 0.3, 0], [-0.4, 0.0, 90], [-0.4, -0.3, 270]]
VSS_YEL = [[0.4, 0.3, 180], [0.4, 0.0, 45], [0.4, -0.3, 0]]


def test_k1_reset_roundtrip_and_units(mk):
    """K1: reset(frame) then get_state returns the same ball (x,y,vx,vy) and robot (x,y,theta deg),
    robot velocities zero (rsim.py:36-38,52-75; vss_gym_base.py:98-103; Frame.py:8)."""
    s = mk(0, 0, 3, 3)
    s.reset([0.1, -0.2, 0.3, -0.4], VSS_BLUE, VSS_YEL)
    st = s.state()[0]
    assert np.allclose(st[:5], [0.1, -0.2, s.field()["ball_radius"], 0.3, -0.4], atol=1e-6)
    exp = np.array(VSS_BLUE + VSS_YEL, dtype=float)
    exp[:, 2] = (exp[:, 2] + 180) % 360 - 180          # reported in (-180, 180]
    for r in range(6):
        row = st[5 + 6 * r:11 + 6 * r]
        assert np.allclose(row[:2], exp[r, :2], atol=1e-6)
        assert abs((row[2] - exp[r, 2] + 180) % 360 - 180) < 1e-4
        assert np.allclose(row[3:], 0.0)
    assert len(st) == 5 + 6 * 6


def test_k1_ssl_state_width_and_field_keys(mk):
    """Frame.py:52-93 (5 + 11 R) and Field.py:4-21 (exactly the 17 keys)."""
    s = mk(1, 2, 1, 6)
    assert s.state().shape[1] == 5 + 11 * 7
    keys = ["length", "width", "penalty_length", "penalty_width", "goal_width", "goal_depth", "ball_radius",
            "rbt_distance_center_kicker", "rbt_kicker_thickness", "rbt_kicker_width", "rbt_wheel0_angle",
            "rbt_wheel1_angle", "rbt_wheel2_angle", "rbt_wheel3_angle", "rbt_radius", "rbt_wheel_radius",
            "rbt_motor_max_rpm"]
    assert list(s.field().keys()) == keys


def test_k2_zero_command_stays_put(mk):
    """K2: zero command => robot stays below 0.05 m/s (dribbling.py:143-145, contested_possession.py:166-167)."""
    for kind, ft, nb, ny, blue, yel in ((0, 0, 3, 3, VSS_BLUE, VSS_YEL), (1, 2, 1, 1, [[0, 0, 0]], [[1.5, 0.2, 180]])):
        s = mk(kind, ft, nb, ny)
        s.reset([0.0, 0.6 if kind == 0 else 1.5, 0, 0], blue, yel)
        before = s.state()
        s.step(np.zeros((s.R, s.C)), k=40)
        after = s.state()
        assert np.abs(after - before).max() < 1e-6


def test_k3_vss_equal_wheels_forward(mk):
    """K3: equal wheel speeds w => steady forward speed w * r_wheel along theta, yaw rate 0; top speed max_v
    (vss_gym_base.py:55-56; vss_gym.py:235-254)."""
    s = mk(0, 0, 3, 3)
    f = s.field()
    s.reset([0, 0.6, 0, 0], [[-0.6, 0.0, 30], [-0.4, 0.5, 90], [-0.4, -0.5, 270]], VSS_YEL)
    cmd = np.zeros((6, 2)); cmd[0] = [20.0, 20.0]
    s.step(cmd, k=12)
    r0 = s.robot(0)[0]
    v = 20.0 * f["rbt_wheel_radius"]
    assert abs(r0[3] - v * np.cos(np.deg2rad(30))) < 1e-4 and abs(r0[4] - v * np.sin(np.deg2rad(30))) < 1e-4
    assert abs(r0[5]) < 1e-3 and abs(r0[2] - 30) < 1e-3
    # saturation at the motor limit: command above max rad/s gives max_v
    s.reset([0, 0.6, 0, 0], [[-0.6, 0.0, 0], [-0.4, 0.5, 90], [-0.4, -0.5, 270]], VSS_YEL)
    cmd[0] = [500.0, 500.0]
    s.step(cmd, k=12)
    max_v = f["rbt_motor_max_rpm"] / 60 * 2 * np.pi * f["rbt_wheel_radius"]
    assert abs(s.robot(0)[0][3] - max_v) < 1e-4


def test_k4_vss_opposite_wheels_spin(mk):
    """K4: opposite wheel speeds +-w => spin rate w * r_wheel / 0.04 rad/s, no translation (vss_gym_base.py:57-58)."""
    s = mk(0, 0, 3, 3)
    f = s.field()
    s.reset([0, 0.6, 0, 0], VSS_BLUE, VSS_YEL)
    cmd = np.zeros((6, 2)); cmd[1] = [-15.0, 15.0]
    s.step(cmd, k=10)
    r1 = s.robot(1)[0]
    assert abs(r1[5] - np.rad2deg(15.0 * f["rbt_wheel_radius"] / 0.04)) < 0.05
    assert abs(r1[3]) < 1e-5 and abs(r1[4]) < 1e-5 and abs(r1[0] + 0.4) < 1e-5


def test_k5_ssl_local_velocity_command(mk):
    """K5: local (vx, vy, vtheta) => body velocity R(theta) (vx, vy), yaw rate vtheta; v_x, v_y reported in
    the field frame; wheel speeds reported and non-zero (static_defenders.py:132-148, 311-322)."""
    s = mk(1, 2, 1, 1)
    th = 60.0
    s.reset([2.0, 1.0, 0, 0], [[0, 0, th]], [[2.5, -1.0, 180]])
    cmd = np.zeros((2, 8)); cmd[0, 1:4] = [1.0, 0.5, 0.0]
    s.step(cmd, k=20)
    r0 = s.robot(0)[0]
    c, sn = np.cos(np.deg2rad(th)), np.sin(np.deg2rad(th))
    assert abs(r0[3] - (c * 1.0 - sn * 0.5)) < 1e-3 and abs(r0[4] - (sn * 1.0 + c * 0.5)) < 1e-3
    assert np.abs(r0[7:11]).min() > 1.0
    cmd[0, 1:4] = [0.0, 0.0, 3.0]
    s.step(cmd, k=30)
    assert abs(s.robot(0)[0][5] - np.rad2deg(3.0)) < 0.1
    # direct wheel-speed mode (rsim.py:137-145): all four wheels at +w spin the robot, no translation
    cmd[0] = [1.0, 30.0, 30.0, 30.0, 30.0, 0, 0, 0]
    s.step(cmd, k=30)
    r0 = s.robot(0)[0]
    assert abs(np.deg2rad(r0[5]) - 30.0 * s.field()["rbt_wheel_radius"] / s.field()["rbt_radius"]) < 1e-2
    assert abs(r0[3]) < 1e-3 and abs(r0[4]) < 1e-3


def test_k6_k7_kick_and_infrared(mk):
    """K7: infrared true with the ball 0.1 m ahead of a robot facing it, false elsewhere
    (dribbling.py:193-195; contested_possession.py:224-225).  K6: ball in the mouth + kick_v_x = 5 =>
    ball departs along the heading at ~5 m/s (static_defenders.py:78, 125)."""
    s = mk(1, 2, 1, 1)
    th = 180.0
    s.reset([-0.1 + 1.0, 0.5, 0, 0], [[1.0, 0.5, th]], [[2.0, -1.0, 0.0]])
    assert s.robot(0)[0][6] == 1.0 and s.robot(1)[0][6] == 0.0
    s.step(np.zeros((2, 8)), k=5)
    assert np.allclose(s.ball()[0][[0, 1]], [0.9, 0.5], atol=1e-6), "ball in the mouth must not be ejected"
    cmd = np.zeros((2, 8)); cmd[0, 5] = 5.0
    s.step(cmd)
    b = s.ball()[0]
    assert b[3] < -4.5 and abs(b[4]) < 1e-3 and b[0] < 0.9 - 0.1
    assert s.robot(0)[0][6] == 0.0
    # ball beside the robot: not touching, kick has no effect
    s.reset([1.0, 0.5 + 0.12, 0, 0], [[1.0, 0.5, th]], [[2.0, -1.0, 0.0]])
    assert s.robot(0)[0][6] == 0.0
    s.step(cmd)
    assert np.abs(s.ball()[0][3:5]).max() < 1e-6


def test_dribbler_holds_ball(mk):
    """A.5: dribbler on + touching => the ball stays with the robot while it translates and rotates
    (dribbling.py:187-202); released when the dribbler is switched off."""
    s = mk(1, 2, 1, 1)
    s.reset([0.1, 0.0, 0, 0], [[0.0, 0.0, 0.0]], [[2.0, -1.0, 0.0]])
    cmd = np.zeros((2, 8)); cmd[0, 1:4] = [-1.0, 0.0, 2.0]; cmd[0, 7] = 1.0
    s.step(cmd, k=30)
    st = s.state()[0]
    bx, by, rx, ry, th = st[0], st[1], st[5], st[6], np.deg2rad(st[7])
    lx = np.cos(th) * (bx - rx) + np.sin(th) * (by - ry)
    ly = -np.sin(th) * (bx - rx) + np.cos(th) * (by - ry)
    assert abs(lx - 0.1) < 1e-4 and abs(ly) < 1e-4 and st[5 + 6] == 1.0
    assert np.hypot(rx, ry) > 0.3
    cmd[0, 7] = 0.0
    s.step(cmd, k=30)
    st = s.state()[0]
    assert np.hypot(st[0] - st[5], st[1] - st[6]) > 0.2 and st[5 + 6] == 0.0


def test_k8_vss_walls_and_goal_mouth(mk):
    """K8: in VSS the ball crosses |x| > L/2 only through a goal mouth (vss_gym.py:161-170)."""
    s = mk(0, 0, 3, 3)
    f = s.field()
    hl = f["length"] / 2
    s.reset([0.6, 0.4, 2.0, 0.0], VSS_BLUE, [[0.4, -0.3, 180], [0.2, -0.5, 45], [0.4, -0.55, 0]])
    xs = []
    for _ in range(40):
        s.step(np.zeros((6, 2)))
        xs.append(s.ball()[0][0])
    assert max(xs) <= hl - f["ball_radius"] + 1e-6          # bounced off the end wall
    assert s.ball()[0][3] < 0 or abs(s.ball()[0][3]) < 1e-3
    s.reset([0.6, 0.05, 2.0, 0.0], VSS_BLUE, [[0.4, -0.3, 180], [0.2, -0.5, 45], [0.4, -0.55, 0]])
    xs = []
    for _ in range(12):
        s.step(np.zeros((6, 2)))
        xs.append(s.ball()[0][0])
    assert max(xs) > hl                                     # entered the goal
    assert max(xs) <= hl + f["goal_depth"] - f["ball_radius"] + 1e-6


def test_k8_ssl_ball_leaves_over_any_line(mk):
    """K8: in SSL the ball can leave the field over any line (static_defenders.py:187-197)."""
    s = mk(1, 2, 1, 1)
    f = s.field()
    s.reset([0.5, 1.8, 0.0, 3.0], [[0, 0, 0]], [[2.0, -1.0, 0.0]])
    s.step(np.zeros((2, 8)), k=10)
    assert s.ball()[0][1] > f["width"] / 2


def test_k9_ball_friction_monotone(mk):
    """K9: a free ball decelerates monotonically to rest and never gains energy."""
    s = mk(0, 0, 3, 3)
    s.reset([-0.5, 0.55, 0.8, 0.0], VSS_BLUE, VSS_YEL)
    sp = []
    for _ in range(80):
        s.step(np.zeros((6, 2)))
        sp.append(np.hypot(*s.ball()[0][3:5]))
    assert all(b <= a + 1e-9 for a, b in zip(sp, sp[1:]))
    assert sp[-1] == 0.0
    decel = (sp[0] - sp[10]) / (10 * 0.025)
    assert abs(decel - 0.05 * 9.81) < 1e-3


def test_k10_ramming_moves_uncommanded_robot(mk):
    """K10: a robot rammed by another exceeds 0.1 m/s (contested_possession.py:165-169)."""
    s = mk(1, 2, 1, 1)
    s.reset([3.0, 1.5, 0, 0], [[0.0, 0.0, 0.0]], [[0.5, 0.0, 180.0]])
    cmd = np.zeros((2, 8)); cmd[0, 1] = 2.0
    vmax = 0.0
    for _ in range(20):
        s.step(cmd)
        vmax = max(vmax, abs(s.robot(1)[0][3]))
    assert vmax > 0.1


def test_k11_k12_row_order_and_non_sticky_commands(mk):
    """K11: rows are blue ids first then yellow at n_blue + id (rsim.py:96-99; Frame.py:28-49).
    K12: commands are not sticky: a robot with a zero row brakes (rsim.py:92-93, 129-130)."""
    s = mk(0, 0, 3, 3)
    s.reset([0, 0.6, 0, 0], VSS_BLUE, VSS_YEL)
    cmd = np.zeros((6, 2)); cmd[3 + 1] = [10.0, 10.0]           # yellow id 1
    s.step(cmd, k=4)
    moved = [np.hypot(*s.robot(r)[0][3:5]) > 0.05 for r in range(6)]
    assert moved == [False, False, False, False, True, False]
    s.step(np.zeros((6, 2)), k=20)
    assert np.hypot(*s.robot(4)[0][3:5]) < 1e-6


def test_batched_envs_are_independent(mk):
    """env i of a batch evolves exactly like the same env alone."""
    n = 5
    rng = np.random.default_rng(0)
    ball = rng.uniform(-0.3, 0.3, (n, 4))
    blue = np.tile(np.array(VSS_BLUE, float), (n, 1, 1)); blue[:, :, 2] = rng.uniform(0, 360, (n, 3))
    yel = np.tile(np.array(VSS_YEL, float), (n, 1, 1))
    cmds = rng.uniform(-30, 30, (n, 6, 2))
    big = mk(0, 0, 3, 3, n)
    big.reset(ball, blue, yel)
    big.step(cmds, k=30)
    for i in range(n):
        one = mk(0, 0, 3, 3, 1)
        one.reset(ball[i], blue[i], yel[i])
        one.step(cmds[i], k=30)
        assert np.array_equal(one.state()[0], big.state()[i])

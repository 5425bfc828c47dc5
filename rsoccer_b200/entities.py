"""Batched mirror of rsoccer_gym/Entities (Ball.py:4-10, Robot.py:5-23, Field.py:4-21,
Frame.py:7-93): same class and field names, same wire layout, but every state attribute
may be a torch tensor of shape [N] (one value per match) instead of a Python float.

`FrameVSS.parse` / `FrameSSL.parse` accept the reference's flat sequence (N = 1, attributes
become scalars exactly as in the reference) or a [N, 5 + K R] tensor from
`BatchedWorld.get_state()` (attributes become zero-copy column views).  Units as in the
reference (Frame.py:8): m, m/s, degrees, degrees/s; field-centred.
"""
from dataclasses import dataclass
from typing import Any, Dict


@dataclass()
class Ball:
    x: Any = None
    y: Any = None
    z: Any = None
    v_x: Any = 0.0
    v_y: Any = 0.0
    v_z: Any = 0.0


@dataclass()
class Robot:
    yellow: bool = None
    id: int = None
    x: Any = None
    y: Any = None
    z: Any = None
    theta: Any = None
    v_x: Any = 0
    v_y: Any = 0
    v_theta: Any = 0
    kick_v_x: Any = 0
    kick_v_z: Any = 0
    dribbler: Any = False
    infrared: Any = False
    wheel_speed: Any = False
    v_wheel0: Any = 0  # rad/s
    v_wheel1: Any = 0  # rad/s
    v_wheel2: Any = 0  # rad/s
    v_wheel3: Any = 0  # rad/s


@dataclass()
class Field:
    length: float
    width: float
    penalty_length: float
    penalty_width: float
    goal_width: float
    goal_depth: float
    ball_radius: float
    rbt_distance_center_kicker: float
    rbt_kicker_thickness: float
    rbt_kicker_width: float
    rbt_wheel0_angle: float
    rbt_wheel1_angle: float
    rbt_wheel2_angle: float
    rbt_wheel3_angle: float
    rbt_radius: float
    rbt_wheel_radius: float
    rbt_motor_max_rpm: float


def _col(state, k):
    """state[k] for a flat sequence, state[:, k] (a view) for a [N, D] tensor / array."""
    if hasattr(state, "ndim") and state.ndim == 2:
        return state[:, k]
    return state[k]


def _as_bool(v):
    return (v != 0) if hasattr(v, "ndim") and getattr(v, "ndim", 0) > 0 else bool(v)


class Frame:
    """Units: seconds, m, m/s, degrees, degrees/s. Reference is field center."""

    def __init__(self):
        self.ball: Ball = Ball()
        self.robots_blue: Dict[int, Robot] = {}
        self.robots_yellow: Dict[int, Robot] = {}


class FrameVSS(Frame):
    ROBOT_WIDTH = 6

    def parse(self, state, n_blues=3, n_yellows=3):
        """Frame.py:17-49: [ball x y z vx vy | per robot x y theta vx vy vtheta], blue first."""
        self.ball.x = _col(state, 0)
        self.ball.y = _col(state, 1)
        self.ball.z = _col(state, 2)
        self.ball.v_x = _col(state, 3)
        self.ball.v_y = _col(state, 4)
        K = self.ROBOT_WIDTH
        for team, n, base in ((self.robots_blue, n_blues, 5), (self.robots_yellow, n_yellows, 5 + n_blues * K)):
            for i in range(n):
                o = base + K * i
                robot = Robot()
                robot.id = i
                robot.x = _col(state, o + 0)
                robot.y = _col(state, o + 1)
                robot.theta = _col(state, o + 2)
                robot.v_x = _col(state, o + 3)
                robot.v_y = _col(state, o + 4)
                robot.v_theta = _col(state, o + 5)
                team[robot.id] = robot
        return self


class FrameSSL(Frame):
    ROBOT_WIDTH = 11

    def parse(self, state, n_blues=3, n_yellows=3):
        """Frame.py:52-93: VSS columns + infrared, v_wheel0..3 per robot."""
        self.ball.x = _col(state, 0)
        self.ball.y = _col(state, 1)
        self.ball.z = _col(state, 2)
        self.ball.v_x = _col(state, 3)
        self.ball.v_y = _col(state, 4)
        K = self.ROBOT_WIDTH
        for team, n, base in ((self.robots_blue, n_blues, 5), (self.robots_yellow, n_yellows, 5 + n_blues * K)):
            for i in range(n):
                o = base + K * i
                robot = Robot()
                robot.id = i
                robot.x = _col(state, o + 0)
                robot.y = _col(state, o + 1)
                robot.theta = _col(state, o + 2)
                robot.v_x = _col(state, o + 3)
                robot.v_y = _col(state, o + 4)
                robot.v_theta = _col(state, o + 5)
                robot.infrared = _as_bool(_col(state, o + 6))
                robot.v_wheel0 = _col(state, o + 7)
                robot.v_wheel1 = _col(state, o + 8)
                robot.v_wheel2 = _col(state, o + 9)
                robot.v_wheel3 = _col(state, o + 10)
                team[robot.id] = robot
        return self

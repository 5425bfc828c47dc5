"""Batched mirror of rsoccer_gym/Simulators/rsim.py: RSim, RSimVSS, RSimSSL.

Same constructor arguments (+ n_envs), same five methods with the same meaning
(`reset(frame)`, `stop()`, `send_commands(commands)`, `get_frame()`, `get_field_params()`),
same row packing (blue ids first, yellow at n_blue + id; unlisted robots get all-zero rows:
rsim.py:92-99, 129-136) and the same error on an out-of-range id (IndexError, rsim.py:100).
Every Frame / Robot attribute may be a scalar (broadcast to all matches) or a [N] tensor.
"""
from typing import List

import torch

from . import engine as _E
from .entities import Field, Frame, FrameSSL, FrameVSS, Robot


class RSim:
    KIND = None

    def __init__(self, field_type: int, n_robots_blue: int, n_robots_yellow: int, time_step_ms: int,
                 n_envs: int = 1, device=None, seed: int = 0, env_offset: int = 0):
        self.n_robots_blue = n_robots_blue
        self.n_robots_yellow = n_robots_yellow
        self.n_envs = n_envs
        # rs_create + rs_bind_state place the robots at the reference's dummy poses (rsim.py:19-24)
        self.simulator = _E.BatchedWorld(self.KIND, field_type, n_robots_blue, n_robots_yellow,
                                         time_step_ms, n_envs, device=device, seed=seed, env_offset=env_offset)
        self.device = self.simulator.device
        self.field = self.get_field_params()

    # rsim.py:36-38
    def reset(self, frame: Frame, mask=None):
        p = self._placement_from_frame(frame)
        self.simulator.reset(p["ball_pos"], p["blue_robots_pos"], p["yellow_robots_pos"], mask=mask)

    # rsim.py:40-41
    def stop(self):
        if self.simulator is not None:
            self.simulator.close()
        self.simulator = None

    def send_commands(self, commands: List[Robot]):
        raise NotImplementedError

    def get_frame(self) -> Frame:
        raise NotImplementedError

    # rsim.py:49-50
    def get_field_params(self) -> Field:
        return Field(**self.simulator.field_params())

    def _vec(self, v):
        """scalar or [N] -> float32 [N] on the device"""
        if torch.is_tensor(v):
            t = v.to(device=self.device, dtype=torch.float32)
            return t.expand(self.n_envs) if t.ndim == 0 else t.reshape(self.n_envs)
        return torch.full((self.n_envs,), float(v), dtype=torch.float32, device=self.device)

    # rsim.py:52-75
    def _placement_from_frame(self, frame: Frame):
        ball = torch.stack([self._vec(frame.ball.x), self._vec(frame.ball.y),
                            self._vec(frame.ball.v_x), self._vec(frame.ball.v_y)], dim=1)

        def team(robots, n):
            if len(robots) != n:
                raise ValueError("frame has %d robots, simulator has %d" % (len(robots), n))
            if n == 0:
                return None
            return torch.stack([torch.stack([self._vec(r.x), self._vec(r.y), self._vec(r.theta)], dim=1)
                                for r in robots.values()], dim=1)
        return {"ball_pos": ball, "blue_robots_pos": team(frame.robots_blue, self.n_robots_blue),
                "yellow_robots_pos": team(frame.robots_yellow, self.n_robots_yellow)}

    def _row(self, cmd: Robot):
        n = self.n_robots_yellow if cmd.yellow else self.n_robots_blue
        if cmd.id is None or not 0 <= cmd.id < n:
            raise IndexError("robot id %r out of range for the %s team" % (cmd.id, "yellow" if cmd.yellow else "blue"))
        return self.n_robots_blue + cmd.id if cmd.yellow else cmd.id


class RSimVSS(RSim):
    KIND = _E.KIND_VSS

    # rsim.py:91-102
    def send_commands(self, commands: List[Robot]):
        R = self.n_robots_blue + self.n_robots_yellow
        sim_commands = torch.zeros(self.n_envs, R, 2, dtype=torch.float32, device=self.device)
        for cmd in commands:
            row = self._row(cmd)
            sim_commands[:, row, 0] = self._vec(cmd.v_wheel0)
            sim_commands[:, row, 1] = self._vec(cmd.v_wheel1)
        self.simulator.step(sim_commands)

    # rsim.py:104-110
    def get_frame(self) -> FrameVSS:
        state = self.simulator.get_state()
        frame = FrameVSS()
        frame.parse(state, self.n_robots_blue, self.n_robots_yellow)
        return frame


class RSimSSL(RSim):
    KIND = _E.KIND_SSL

    # rsim.py:128-155
    def send_commands(self, commands: List[Robot]):
        R = self.n_robots_blue + self.n_robots_yellow
        c = torch.zeros(self.n_envs, R, 8, dtype=torch.float32, device=self.device)
        for cmd in commands:
            row = self._row(cmd)
            ws = cmd.wheel_speed
            if torch.is_tensor(ws):
                raise TypeError("wheel_speed selects the command layout and must be a plain bool")
            c[:, row, 0] = float(bool(ws))
            if ws:
                c[:, row, 1] = self._vec(cmd.v_wheel0)
                c[:, row, 2] = self._vec(cmd.v_wheel1)
                c[:, row, 3] = self._vec(cmd.v_wheel2)
                c[:, row, 4] = self._vec(cmd.v_wheel3)
            else:
                c[:, row, 1] = self._vec(cmd.v_x)
                c[:, row, 2] = self._vec(cmd.v_y)
                c[:, row, 3] = self._vec(cmd.v_theta)
            c[:, row, 5] = self._vec(cmd.kick_v_x)
            c[:, row, 6] = self._vec(cmd.kick_v_z)
            c[:, row, 7] = self._vec(cmd.dribbler)
        self.simulator.step(c)

    # rsim.py:157-163
    def get_frame(self) -> FrameSSL:
        state = self.simulator.get_state()
        frame = FrameSSL()
        frame.parse(state, self.n_robots_blue, self.n_robots_yellow)
        return frame

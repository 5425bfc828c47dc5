"""Batched mirrors of VSSBaseEnv / SSLBaseEnv (rsoccer_gym/vss/vss_gym_base.py,
ssl/ssl_gym_base.py): the same step / reset orchestration (vss_gym_base.py:72-106) and the
same four subclass hooks, over N matches:

    _get_commands(actions)            -> List[Robot]     (attributes scalar or [N] tensors)
    _frame_to_observations()          -> [N, n_obs] tensor
    _calculate_reward_and_done()      -> ([N] reward, [N] done)
    _get_initial_positions_frame()    -> Frame            (attributes scalar or [N] tensors)

`self.frame` / `self.last_frame` are batched Frame views (entities.py).  This is the generic
path for USER subclasses (one launch for physics, torch ops for the hooks); the three
benchmarked envs override `step` with the single fused launch (envs/vss.py, envs/ssl.py).
Rendering is out of scope.
"""
import math
from typing import List

import torch

from ..entities import Frame, Robot
from ..simulators import RSimSSL, RSimVSS


class BoxSpec:
    """shape/low/high record for the batched action / observation spaces"""

    def __init__(self, low, high, shape, dtype=torch.float32):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    def sample(self, generator=None, device=None):
        return torch.rand(self.shape, generator=generator, device=device) * (self.high - self.low) + self.low


class _BaseVecEnv:
    NORM_BOUNDS = 1.2
    RSIM = None
    HALF_AXLE = None     # 0.04 (VSS) / 0.095 (SSL): vss_gym_base.py:57-58, ssl_gym_base.py:58-59

    def __init__(self, field_type: int, n_robots_blue: int, n_robots_yellow: int, time_step: float,
                 num_envs: int = 1, device=None, seed: int = 0, env_offset: int = 0, render_mode=None):
        if render_mode not in (None, "rgb_array"):
            raise NotImplementedError("render_mode %r: only 'rgb_array' (no window) is offered" % (render_mode,))
        self.render_mode = render_mode
        self.num_envs = num_envs
        self.time_step = time_step
        self.rsim = self.RSIM(field_type=field_type, n_robots_blue=n_robots_blue,
                              n_robots_yellow=n_robots_yellow, time_step_ms=int(self.time_step * 1000),
                              n_envs=num_envs, device=device, seed=seed, env_offset=env_offset)
        self.device = self.rsim.device
        self.n_robots_blue = n_robots_blue
        self.n_robots_yellow = n_robots_yellow
        self.field_type = field_type
        self.field = self.rsim.get_field_params()
        self.max_pos = max(self.field.width / 2, (self.field.length / 2) + self.field.penalty_length)
        max_wheel_rad_s = (self.field.rbt_motor_max_rpm / 60) * 2 * math.pi
        self.max_v = max_wheel_rad_s * self.field.rbt_wheel_radius
        self.max_w = math.degrees(self.max_v / self.HALF_AXLE)
        self.frame: Frame = None
        self.last_frame: Frame = None
        self.steps = 0
        self.sent_commands = None

    # vss_gym_base.py:72-90
    def step(self, action):
        self.steps += 1
        commands: List[Robot] = self._get_commands(action)
        self.rsim.send_commands(commands)
        self.sent_commands = commands
        self.last_frame = self.frame
        self.frame = self.rsim.get_frame()
        observation = self._frame_to_observations()
        reward, done = self._calculate_reward_and_done()
        return observation, reward, done, torch.zeros_like(done, dtype=torch.bool), {}

    # vss_gym_base.py:92-106
    def reset(self, *, seed=None, options=None):
        self.steps = 0
        self.last_frame = None
        self.sent_commands = None
        initial_pos_frame: Frame = self._get_initial_positions_frame()
        self.rsim.reset(initial_pos_frame)
        self.frame = self.rsim.get_frame()
        return self._frame_to_observations(), {}

    def close(self):
        self.rsim.stop()

    def render(self, index=0, width_px=750):
        """RGB picture [H, W, 3] uint8 of match `index` (vss_gym_base.py:148-187, rgb_array mode)."""
        from ..render import render_rgb
        row = self.rsim.simulator.get_state()[int(index)].cpu().numpy()
        return render_rgb(row, self.rsim.simulator.field_params(), "vss" if isinstance(self, VSSBaseVecEnv) else "ssl",
                          self.n_robots_blue, self.n_robots_yellow, width_px=width_px)

    def _get_commands(self, action):
        raise NotImplementedError

    def _frame_to_observations(self):
        raise NotImplementedError

    def _calculate_reward_and_done(self):
        raise NotImplementedError

    def _get_initial_positions_frame(self) -> Frame:
        raise NotImplementedError

    # vss_gym_base.py:213-220
    def norm_pos(self, pos):
        return torch.clamp(pos / self.max_pos, -self.NORM_BOUNDS, self.NORM_BOUNDS)

    def norm_v(self, v):
        return torch.clamp(v / self.max_v, -self.NORM_BOUNDS, self.NORM_BOUNDS)

    def norm_w(self, w):
        return torch.clamp(w / self.max_w, -self.NORM_BOUNDS, self.NORM_BOUNDS)


class VSSBaseVecEnv(_BaseVecEnv):
    RSIM = RSimVSS
    HALF_AXLE = 0.04


class SSLBaseVecEnv(_BaseVecEnv):
    RSIM = RSimSSL
    HALF_AXLE = 0.095

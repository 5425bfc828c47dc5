"""The five registered reference envs as batched envs whose `step` is ONE fused launch.

    VSSVecEnv                       <- rsoccer_gym/vss/env_vss/vss_gym.py           (VSS-v0)
    SSLStaticDefendersVecEnv        <- ssl/ssl_hw_challenge/static_defenders.py     (SSLStaticDefenders-v0)
    SSLContestedPossessionVecEnv    <- ssl/ssl_hw_challenge/contested_possession.py (SSLContestedPossession-v0)
    SSLDribblingVecEnv              <- ssl/ssl_hw_challenge/dribbling.py            (SSLDribbling-v0)
    SSLPassEnduranceVecEnv          <- ssl/ssl_hw_challenge/pass_endurance.py       (SSLPassEndurance-v0)

API (gymnasium VectorEnv flavour, same-step auto-reset):
    obs, info = env.reset()
    obs, reward, terminated, truncated, info = env.step(actions)     # torch CUDA tensors, [N, ...]
`info` is the reference's `reward_shaping_total` dict (vss_gym.py:150-158) with one [N] tensor
per key (zero-copy views of the on-device accumulators).  When an episode ends the returned
observation is the first one of the next episode; `truncated` follows the registered step
limits (rsoccer_gym/__init__.py:3-30).

Terminal observations: construct the env with `final_obs=True` and `info["final_obs"]` holds every env's observation
BEFORE any reset of this step (the terminal observation of the envs flagged in `info["_final_obs"]` = terminated |
truncated: what a time-limit bootstrap needs, gymnasium's same-step autoreset convention).  That mode steps without
the in-kernel reset and re-places the ended matches with one masked `rs_task_reset` launch (three launches per
step instead of one; placements then come from the reset stream of the Philox generator instead of the auto-reset one).

Aliasing: `step` / `reset` return the SAME preallocated tensors on every call (zero-copy: the
kernel writes them in place).  A rollout buffer must copy what it keeps (`buf[t].copy_(obs)`),
or construct the env with `copy_outputs=True` to get fresh tensors from every call.
"""
import torch

from .. import engine as _E
from ..entities import FrameSSL, FrameVSS
from .base import BoxSpec

VSS_INFO_KEYS = ("goal_score", "move", "ball_grad", "energy", "goals_blue", "goals_yellow")
SSL_INFO_KEYS = ("goal", "rbt_in_gk_area", "done_ball_out", "done_ball_out_right", "done_rbt_out",
                 "ball_dist", "ball_grad", "energy", "collision")


class _FusedVecEnv:
    KIND = None
    TASK = None
    FIELD_TYPE = None
    N_BLUE = N_YELLOW = None
    ACT_DIM = None
    MAX_EPISODE_STEPS = None
    INFO_KEYS = ()
    NORM_BOUNDS = 1.2

    def __init__(self, num_envs=1, device=None, seed=0, env_offset=0, auto_reset=True, max_episode_steps=None,
                 field_type=None, render_mode=None, copy_outputs=False, final_obs=False):
        if render_mode not in (None, "rgb_array"):
            raise NotImplementedError("render_mode %r: only 'rgb_array' (no window) is offered" % (render_mode,))
        self.render_mode = render_mode
        self.num_envs = int(num_envs)
        self.auto_reset = bool(auto_reset)
        self.max_episode_steps = int(max_episode_steps or self.MAX_EPISODE_STEPS)
        self.time_step = 0.025
        ft = self.FIELD_TYPE if field_type is None else field_type
        self.world = _E.BatchedWorld(self.KIND, ft, self.N_BLUE, self.N_YELLOW, 25, self.num_envs, device=device,
                                     seed=seed, env_offset=env_offset)
        self.device = self.world.device
        self.n_robots_blue, self.n_robots_yellow = self.N_BLUE, self.N_YELLOW
        self.obs_dim = self.world.obs_dim(self.TASK)
        self.action_space = BoxSpec(-1.0, 1.0, (self.num_envs, self.ACT_DIM))
        self.observation_space = BoxSpec(-self.NORM_BOUNDS, self.NORM_BOUNDS, (self.num_envs, self.obs_dim))
        self._out = self.world.alloc_outputs(self.TASK)
        self.copy_outputs = bool(copy_outputs)
        self.final_obs = bool(final_obs)
        self._final = torch.empty_like(self._out[0]) if self.final_obs else None
        self._pinned = None
        self.field = None

    # ---- gym surface
    def reset(self, *, seed=None, options=None, mask=None):
        # a masked reset rewrites only the rows of the selected matches: the others keep their
        # current observation (the one the last step wrote into the same tensor)
        obs = self.world.task_reset(self.TASK, mask=mask, obs=self._out[0])
        return (obs.clone() if self.copy_outputs else obs), self.info()

    def step(self, actions):
        if self.final_obs and self.auto_reset:
            return self._step_with_final_obs(actions)
        obs, rew, done, trunc = self._step(actions)
        if self.copy_outputs:
            obs, rew = obs.clone(), rew.clone()
        return obs, rew, done.bool(), trunc.bool(), self.info()

    def _step_with_final_obs(self, actions):
        # the fused step without its in-kernel reset: the observation rows of ended matches are terminal observations
        obs, rew, done, trunc = self._step(actions, auto_reset=False)
        ended = done | trunc
        self._final.copy_(obs)
        self.world.task_reset(self.TASK, mask=ended, obs=obs)      # rewrites the rows (and the state) of the ended matches only
        info = self.info()
        info["final_obs"] = self._final.clone() if self.copy_outputs else self._final
        info["_final_obs"] = ended.bool()
        if self.copy_outputs:
            obs, rew = obs.clone(), rew.clone()
        return obs, rew, done.bool(), trunc.bool(), info

    def step_raw(self, actions):
        """step without building the info dict / bool casts: (obs, reward, done u8, trunc u8)"""
        return self._step(actions)

    def info(self):
        n = self.num_envs
        return {k: self.world.info[i, :n] for i, k in enumerate(self.INFO_KEYS)}

    def close(self):
        self.world.close()

    # ---- host-buffer end-to-end call (numpy in / numpy out through pinned memory)
    def step_host(self, actions_np):
        if self._pinned is None:
            n = self.num_envs
            self._pinned = (torch.empty(n, self.ACT_DIM, dtype=torch.float32).pin_memory(),
                            ) + self.world.alloc_host_outputs(self.TASK)      # one block -> one D2H copy
        a, o, r, d, t = self._pinned
        a.copy_(torch.as_tensor(actions_np, dtype=torch.float32).reshape(a.shape))
        self._step_host(a, o, r, d, t)
        return o.numpy(), r.numpy(), d.numpy().astype(bool), t.numpy().astype(bool)

    # ---- the same, split-phase (gymnasium VectorEnv's step_async / step_wait): with two env groups on two CUDA
    # streams one group's observations cross PCIe while the other group steps
    def step_async(self, actions_np):
        if self._pinned is None:
            n = self.num_envs
            self._pinned = (torch.empty(n, self.ACT_DIM, dtype=torch.float32).pin_memory(),
                            ) + self.world.alloc_host_outputs(self.TASK)
        a, o, r, d, t = self._pinned
        a.copy_(torch.as_tensor(actions_np, dtype=torch.float32).reshape(a.shape))
        self._step_host(a, o, r, d, t, begin=True)

    def step_wait(self):
        self.world.host_step_wait()
        _, o, r, d, t = self._pinned
        return o.numpy(), r.numpy(), d.numpy().astype(bool), t.numpy().astype(bool)

    def render(self, index=0, width_px=750):
        """RGB picture [H, W, 3] uint8 of match `index` (vss_gym_base.py:148-187, rgb_array mode)."""
        from ..render import render_rgb
        row = self.world.get_state()[int(index)].cpu().numpy()
        return render_rgb(row, self.world.field_params(), "vss" if self.KIND == _E.KIND_VSS else "ssl",
                          self.N_BLUE, self.N_YELLOW, width_px=width_px)

    # ---- batched Frame view of the current state (Entities/Frame.py layout)
    @property
    def frame(self):
        st = self.world.get_state()
        f = FrameVSS() if self.KIND == _E.KIND_VSS else FrameSSL()
        return f.parse(st, self.N_BLUE, self.N_YELLOW)


class VSSVecEnv(_FusedVecEnv):
    KIND, TASK, FIELD_TYPE, N_BLUE, N_YELLOW, ACT_DIM = _E.KIND_VSS, _E.TASK_VSS_V0, 0, 3, 3, 2
    MAX_EPISODE_STEPS = 1200              # rsoccer_gym/__init__.py:4
    INFO_KEYS = VSS_INFO_KEYS

    def _step(self, actions, auto_reset=None):
        return self.world.vss_env_step(actions, auto_reset=self.auto_reset if auto_reset is None else auto_reset,
                                       max_steps=self.max_episode_steps, out=self._out)

    def _step_host(self, a, o, r, d, t, begin=False):
        f = self.world.vss_env_step_host_begin if begin else self.world.vss_env_step_host
        f(a, o, r, d, t, auto_reset=self.auto_reset, max_steps=self.max_episode_steps)


class _SSLFused(_FusedVecEnv):
    KIND, FIELD_TYPE, ACT_DIM = _E.KIND_SSL, 2, 5
    INFO_KEYS = SSL_INFO_KEYS

    def _step(self, actions, auto_reset=None):
        return self.world.ssl_env_step(self.TASK, actions, auto_reset=self.auto_reset if auto_reset is None else auto_reset,
                                       max_steps=self.max_episode_steps, out=self._out)

    def _step_host(self, a, o, r, d, t, begin=False):
        f = self.world.ssl_env_step_host_begin if begin else self.world.ssl_env_step_host
        f(self.TASK, a, o, r, d, t, auto_reset=self.auto_reset, max_steps=self.max_episode_steps)


class SSLStaticDefendersVecEnv(_SSLFused):
    TASK, N_BLUE, N_YELLOW = _E.TASK_SSL_STATIC_DEFENDERS_V0, 1, 6
    MAX_EPISODE_STEPS = 1000              # rsoccer_gym/__init__.py:11


class SSLContestedPossessionVecEnv(_SSLFused):
    TASK, N_BLUE, N_YELLOW = _E.TASK_SSL_CONTESTED_POSSESSION_V0, 1, 1
    MAX_EPISODE_STEPS = 1200              # rsoccer_gym/__init__.py:24


class SSLDribblingVecEnv(_SSLFused):
    """SSLDribbling-v0: zig-zag course between four parked robots; `checkpoints` (the reference's
    checkpoints_count, dribbling.py:56) is a zero-copy [N] view of the on-device counter."""
    TASK, N_BLUE, N_YELLOW, ACT_DIM = _E.TASK_SSL_DRIBBLING_V0, 1, 4, 4
    MAX_EPISODE_STEPS = 4800              # rsoccer_gym/__init__.py:17
    INFO_KEYS = ()                        # ssl_gym_base.py:88 returns {} and dribbling.py adds nothing

    @property
    def checkpoints(self):
        return self.world.task_word[:self.num_envs]


class SSLPassEnduranceVecEnv(_SSLFused):
    """SSLPassEndurance-v0: the shooter turns / kicks, the receiver holds its dribbler on;
    info = reward_shaping_total {reversed_dist, ball_grad} (pass_endurance.py:113-114)."""
    TASK, N_BLUE, N_YELLOW, ACT_DIM = _E.TASK_SSL_PASS_ENDURANCE_V0, 2, 0, 3
    MAX_EPISODE_STEPS = 1200              # rsoccer_gym/__init__.py:29
    INFO_KEYS = ("reversed_dist", "ball_grad")

"""Minimal stand-in for `gymnasium` (absent from this image).

Only what the reference package touches: `Env`, `spaces.Box`, `envs.registration.register`,
`make` with entry-point loading + kwargs merge, and the `TimeLimit` wrapper that sets
`truncated` at `max_episode_steps` (rsoccer_gym/__init__.py:3-30).  Installed into
`sys.modules['gymnasium']` by `rsoccer_b200.compat.install()` only when the real package
cannot be imported.
"""
import importlib

import numpy as np

from . import spaces  # noqa: F401
from .envs import registration
from .envs.registration import register, registry  # noqa: F401

__version__ = "0.0-shim"


class Env:
    metadata = {"render_modes": []}
    render_mode = None
    action_space = None
    observation_space = None
    spec = None

    def reset(self, *, seed=None, options=None):
        if seed is not None or not hasattr(self, "np_random"):
            self.np_random = np.random.default_rng(seed)
        return None

    def step(self, action):
        raise NotImplementedError

    def render(self):
        return None

    def close(self):
        pass

    @property
    def unwrapped(self):
        return self


class Wrapper:
    """forwards everything it does not define to the wrapped env (spaces, metadata, hooks)"""

    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        if name.startswith("_") or name == "env":
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def step(self, action):
        return self.env.step(action)

    def render(self):
        return self.env.render()

    def close(self):
        return self.env.close()


class TimeLimit(Wrapper):
    """truncated = True once `max_episode_steps` steps elapsed since reset."""

    def __init__(self, env, max_episode_steps):
        super().__init__(env)
        self._max_episode_steps = max_episode_steps
        self._elapsed_steps = 0

    def reset(self, **kwargs):
        self._elapsed_steps = 0
        return self.env.reset(**kwargs)

    def step(self, action):
        obs, reward, terminated, truncated, info = self.env.step(action)
        self._elapsed_steps += 1
        if self._elapsed_steps >= self._max_episode_steps:
            truncated = True
        return obs, reward, terminated, truncated, info


def make(id, **kwargs):
    spec = registration.registry[id]
    entry = spec.entry_point
    if isinstance(entry, str):
        mod, _, attr = entry.partition(":")
        entry = getattr(importlib.import_module(mod), attr)
    kw = dict(spec.kwargs)
    kw.update(kwargs)
    env = entry(**kw)
    env.spec = spec
    if spec.max_episode_steps is not None:
        env = TimeLimit(env, spec.max_episode_steps)
    return env

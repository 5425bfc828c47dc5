"""Drop-in glue: lets the UNMODIFIED reference package (rsoccer_gym) run on this engine.

`install()` puts a `robosim`-compatible module backed by the CUDA engine into
`sys.modules` (the reference does `import robosim` at Simulators/rsim.py:2) and, only when
they are missing from the interpreter, import-level stand-ins for `gymnasium` and `pygame`.
After that, `import rsoccer_gym; gym.make('VSS-v0')` works as in the reference README
(README.md:116-133) with N = 1.  For N >> 1 use `rsoccer_b200.envs`.
"""
import importlib
import sys


def _missing(name):
    try:
        importlib.import_module(name)
        return False
    except Exception:
        return True


def install(robosim_module=None, force_shims=False):
    """robosim_module: module object to expose as `robosim` (default: the CUDA-backed
    rsoccer_b200.compat.robosim).  Returns the list of module names that were installed."""
    done = []
    if force_shims or _missing("gymnasium"):
        from . import gymnasium_shim
        sys.modules["gymnasium"] = gymnasium_shim
        sys.modules["gymnasium.spaces"] = gymnasium_shim.spaces
        sys.modules["gymnasium.envs"] = gymnasium_shim.envs
        sys.modules["gymnasium.envs.registration"] = gymnasium_shim.envs.registration
        done.append("gymnasium")
    if force_shims or _missing("pygame"):
        from . import pygame_shim
        sys.modules["pygame"] = pygame_shim
        done.append("pygame")
    if robosim_module is None:
        from . import robosim as robosim_module
    sys.modules["robosim"] = robosim_module
    done.append("robosim")
    return done

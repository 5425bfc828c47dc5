"""`robosim` stand-in backed by the CPU oracle -- TEST INFRASTRUCTURE (golden generation,
reference-env smoke tests).  Same surface as the pybind11 module the reference imports at
rsoccer_gym/Simulators/rsim.py:2."""
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
from oracle import oracle as _O  # noqa: E402


class _Sim:
    KIND = None

    def __init__(self, field_type, n_robots_blue, n_robots_yellow, time_step_ms, ball_pos,
                 blue_robots_pos, yellow_robots_pos):
        self._w = _O.OracleWorld(self.KIND, int(field_type), int(n_robots_blue), int(n_robots_yellow),
                                 int(time_step_ms), 1)
        self._nb, self._ny = int(n_robots_blue), int(n_robots_yellow)
        self.reset(ball_pos, blue_robots_pos, yellow_robots_pos)

    def reset(self, ball_pos, blue_robots_pos, yellow_robots_pos):
        self._w.reset(np.asarray(ball_pos, dtype=np.float64).reshape(1, 4),
                      np.asarray(blue_robots_pos, dtype=np.float64).reshape(1, self._nb, 3),
                      np.asarray(yellow_robots_pos, dtype=np.float64).reshape(1, self._ny, 3))

    def step(self, commands):
        self._w.step(np.asarray(commands, dtype=np.float64))

    def get_state(self):
        return self._w.get_state().reshape(-1)

    def get_field_params(self):
        return self._w.field_params()


class VSS(_Sim):
    KIND = _O.KIND_VSS


class SSL(_Sim):
    KIND = _O.KIND_SSL

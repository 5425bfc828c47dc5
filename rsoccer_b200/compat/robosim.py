"""`robosim`-compatible module backed by the CUDA engine (N = 1).

Same constructor and the same five methods the reference calls on the third-party pybind11
classes (rsoccer_gym/Simulators/rsim.py:38, 50, 102, 105, 116-124, 155, 158, 169-177):

    VSS(field_type, n_robots_blue, n_robots_yellow, time_step_ms, ball_pos, blue_robots_pos,
        yellow_robots_pos)          .step(cmds)  .get_state()  .reset(ball, blue, yellow)
    SSL(... same ...)               .get_field_params()

Arrays are float64 on the wire like the reference's (rsim.py:93, 130); the engine computes
in fp32.  One D2H copy of 5 + K*R floats per get_state(): fine for plumbing / parity at
N = 1 (BASELINE config 1), irrelevant to the batched metric.
"""
import numpy as np

from .. import engine as _E


class _Sim:
    KIND = None

    def __init__(self, field_type, n_robots_blue, n_robots_yellow, time_step_ms, ball_pos,
                 blue_robots_pos, yellow_robots_pos):
        self._w = _E.BatchedWorld(self.KIND, int(field_type), int(n_robots_blue), int(n_robots_yellow),
                                  int(time_step_ms), 1)
        self._nb, self._ny = int(n_robots_blue), int(n_robots_yellow)
        self.reset(np.asarray(ball_pos, dtype=np.float64), np.asarray(blue_robots_pos, dtype=np.float64),
                   np.asarray(yellow_robots_pos, dtype=np.float64))

    def reset(self, ball_pos, blue_robots_pos, yellow_robots_pos):
        ball = np.asarray(ball_pos, dtype=np.float32).reshape(1, 4)
        blue = np.asarray(blue_robots_pos, dtype=np.float32).reshape(1, self._nb, 3) if self._nb else None
        yellow = np.asarray(yellow_robots_pos, dtype=np.float32).reshape(1, self._ny, 3) if self._ny else None
        self._w.reset(ball, blue, yellow)

    def step(self, commands):
        c = np.asarray(commands, dtype=np.float32)
        if c.shape != (self._w.R, self._w.cmd_dim):
            raise IndexError("commands must have shape (%d, %d)" % (self._w.R, self._w.cmd_dim))
        self._w.step(c.reshape(1, self._w.R, self._w.cmd_dim))

    def get_state(self):
        return self._w.get_state().cpu().numpy().astype(np.float64).reshape(-1)

    def get_field_params(self):
        return self._w.field_params()


class VSS(_Sim):
    KIND = _E.KIND_VSS


class SSL(_Sim):
    KIND = _E.KIND_SSL

"""ctypes/numpy wrapper of the CPU oracle (oracle/rs_oracle.c).

TEST INFRASTRUCTURE -- not product code.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / ``--impl reference`` legs import this module; the
product package ``rsoccer_b200`` never does.

PARITY UNPINNED for the physics (see the header of rs_oracle.c); the task logic
(commands / observation / reward / done) is pinned by tests/golden.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "librs_oracle.so")

KIND_VSS, KIND_SSL = 0, 1
TASK_VSS, TASK_SSL_STATIC_DEFENDERS, TASK_SSL_CONTESTED_POSSESSION = 0, 1, 2
TASK_SSL_DRIBBLING, TASK_SSL_PASS_ENDURANCE = 3, 4
TASK_ACT = {0: 2, 1: 5, 2: 5, 3: 4, 4: 3}


def task_obs_dim(task, nb, ny):
    return {0: 4 + 7 * nb + 5 * ny, 3: 5 + 8 * nb + 2 * ny, 4: 4 + 6 * nb}.get(task, 4 + 8 * nb + 2 * ny)

INFO_W = 9

FIELD_KEYS = (
    "length", "width", "penalty_length", "penalty_width", "goal_width", "goal_depth",
    "ball_radius", "rbt_distance_center_kicker", "rbt_kicker_thickness", "rbt_kicker_width",
    "rbt_wheel0_angle", "rbt_wheel1_angle", "rbt_wheel2_angle", "rbt_wheel3_angle",
    "rbt_radius", "rbt_wheel_radius", "rbt_motor_max_rpm",
)


def build(force=False):
    """Compile librs_oracle.so next to this file (gcc, seconds)."""
    src = os.path.join(_HERE, "rs_oracle.c")
    spec = os.path.join(_HERE, "..", "include", "rs_spec.h")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(spec))):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, i32, u64, i64 = C.c_void_p, C.c_int, C.c_uint64, C.c_int64
        L.orc_create.restype = vp
        L.orc_create.argtypes = [i32, i32, i32, i32, i32, i32, u64, i64]
        L.orc_destroy.argtypes = [vp]
        L.orc_set_threads.argtypes = [vp, i32]
        L.orc_field_params.argtypes = [vp, vp]
        L.orc_params_dump.argtypes = [vp, vp]
        L.orc_params_dump.restype = i32
        L.orc_get_t.restype = u64
        L.orc_get_t.argtypes = [vp]
        L.orc_set_t.argtypes = [vp, u64]
        L.orc_reset.argtypes = [vp, vp, vp, vp, vp]
        L.orc_set_raw.argtypes = [vp, vp]
        L.orc_get_raw.argtypes = [vp, vp]
        L.orc_step.argtypes = [vp, vp]
        L.orc_get_state.argtypes = [vp, vp]
        L.orc_get_margin.argtypes = [vp, vp]
        L.orc_get_task_state.argtypes = [vp, vp, vp, vp, vp, vp]
        L.orc_set_task_state.argtypes = [vp, vp, vp, vp, vp, vp]
        L.orc_task_reset.argtypes = [vp, i32, vp]
        L.orc_vss_env_step.argtypes = [vp, vp, vp, i32, i32, vp, vp, vp, vp, vp]
        L.orc_ssl_env_step.argtypes = [vp, i32, vp, i32, i32, vp, vp, vp, vp, vp]
        L.orc_ssl_hw_env_step.argtypes = [vp, i32, vp, i32, i32, vp, vp, vp, vp, vp]
        L.orc_task_obs.argtypes = [vp, i32, vp]
        L.orc_philox.argtypes = [vp, vp, vp]
        L.orc_max_threads.restype = i32
        L.orc_has_openmp.restype = i32
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def philox4x32_10(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox(_p(c), _p(k), _p(out))
    return out


def max_threads():
    return int(lib().orc_max_threads())


def usable_threads(want):
    """threads the oracle can really use: `want` when built with OpenMP (its pragma's num_threads()
    overrides OMP_NUM_THREADS), else 1"""
    return max(1, int(want)) if lib().orc_has_openmp() else 1


class OracleWorld:
    """N independent matches stepped by the scalar fp64 oracle."""

    def __init__(self, kind, field_type, n_blue, n_yellow, time_step_ms=25, n_envs=1, seed=0,
                 env_offset=0, threads=1):
        self.L = lib()
        self.h = self.L.orc_create(kind, field_type, n_blue, n_yellow, time_step_ms, n_envs,
                                   seed, env_offset)
        if not self.h:
            raise ValueError("orc_create failed (unknown world or bad sizes)")
        self.kind, self.n, self.nb, self.ny = kind, n_envs, n_blue, n_yellow
        self.R = n_blue + n_yellow
        self.K = 6 if kind == KIND_VSS else 11
        self.Ccmd = 2 if kind == KIND_VSS else 8
        self.L.orc_set_threads(self.h, threads)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_destroy(self.h)
            self.h = None

    def set_threads(self, n):
        self.L.orc_set_threads(self.h, n)

    def field_params(self):
        out = np.zeros(17)
        self.L.orc_field_params(self.h, _p(out))
        return dict(zip(FIELD_KEYS, out.tolist()))

    def params(self):
        """the derived parameter block of include/rs_spec.h (walls, masses, drive limits, omni matrices, kicker box)"""
        out = np.zeros(64)
        n = self.L.orc_params_dump(self.h, _p(out))
        assert n == 56, n
        d = dict(dt=out[0], h=out[1], x_out=out[2], y_out=out[3], x_near=out[4], n_box=int(out[5]),
                 box=out[6:14].reshape(2, 4).copy())
        names = ("ball_mass", "rbt_mass", "e_ball_wall", "e_rbt_wall", "e_ball_rbt", "e_rbt_rbt", "mu_ball_rbt",
                 "ball_decel", "wheel_max_rad_s", "half_track", "acc_fwd", "acc_lat", "acc_ang")
        d.update(zip(names, out[14:27].tolist()))
        d["omni_J"] = out[27:39].reshape(4, 3).copy()
        d["omni_Jpinv"] = out[39:51].reshape(3, 4).copy()
        d.update(zip(("kick_centre", "kick_reach", "kick_half_width", "mouth_half_chord", "kick_speed_max"),
                     out[51:56].tolist()))
        return d

    @property
    def t(self):
        return int(self.L.orc_get_t(self.h))

    @t.setter
    def t(self, v):
        self.L.orc_set_t(self.h, int(v))

    def reset(self, ball, blue, yellow, mask=None):
        ball = np.ascontiguousarray(ball, dtype=np.float64).reshape(self.n, 4)
        blue = np.ascontiguousarray(blue, dtype=np.float64).reshape(self.n, self.nb, 3)
        yellow = np.ascontiguousarray(yellow, dtype=np.float64).reshape(self.n, self.ny, 3)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self.L.orc_reset(self.h, _p(ball), _p(blue), _p(yellow), _p(m))

    def set_raw(self, s):
        s = np.ascontiguousarray(s, dtype=np.float64).reshape(self.n, 4 + 6 * self.R)
        self.L.orc_set_raw(self.h, _p(s))

    def get_raw(self):
        s = np.zeros((self.n, 4 + 6 * self.R))
        self.L.orc_get_raw(self.h, _p(s))
        return s

    def step(self, cmds):
        cmds = np.ascontiguousarray(cmds, dtype=np.float64).reshape(self.n, self.R, self.Ccmd)
        self.L.orc_step(self.h, _p(cmds))

    def get_state(self):
        out = np.zeros((self.n, 5 + self.K * self.R))
        self.L.orc_get_state(self.h, _p(out))
        return out

    def margin(self):
        out = np.zeros(self.n)
        self.L.orc_get_margin(self.h, _p(out))
        return out

    def get_task_state(self):
        ou = np.zeros((self.n, 2 * (self.R - 1)))
        pp = np.zeros(self.n)
        hp = np.zeros(self.n, dtype=np.int32)
        st = np.zeros(self.n, dtype=np.int32)
        info = np.zeros((self.n, INFO_W))
        self.L.orc_get_task_state(self.h, _p(ou), _p(pp), _p(hp), _p(st), _p(info))
        return dict(ou=ou, prev_pot=pp, has_prev=hp, steps=st, info=info)

    def set_task_state(self, ou=None, prev_pot=None, has_prev=None, steps=None, info=None):
        def c(a, dt):
            return None if a is None else np.ascontiguousarray(a, dtype=dt)
        a, b, d, e, f = c(ou, np.float64), c(prev_pot, np.float64), c(has_prev, np.int32), \
            c(steps, np.int32), c(info, np.float64)
        self.L.orc_set_task_state(self.h, _p(a), _p(b), _p(d), _p(e), _p(f))

    def task_reset(self, task, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self.L.orc_task_reset(self.h, task, _p(m))

    def task_obs(self, task):
        obs = np.zeros((self.n, task_obs_dim(task, self.nb, self.ny)))
        self.L.orc_task_obs(self.h, task, _p(obs))
        return obs

    def vss_env_step(self, actions, normals=None, auto_reset=True, max_steps=1200, want_cmds=False):
        a = np.ascontiguousarray(actions, dtype=np.float32).reshape(self.n, 2)
        z = None if normals is None else np.ascontiguousarray(normals, dtype=np.float64).reshape(
            self.n, 2 * (self.R - 1))
        obs = np.zeros((self.n, 4 + 7 * self.nb + 5 * self.ny))
        rew = np.zeros(self.n)
        done = np.zeros(self.n, dtype=np.uint8)
        trunc = np.zeros(self.n, dtype=np.uint8)
        cmds = np.zeros((self.n, self.R, 2)) if want_cmds else None
        self.L.orc_vss_env_step(self.h, _p(a), _p(z), int(auto_reset), max_steps, _p(obs), _p(rew),
                                _p(done), _p(trunc), _p(cmds))
        return (obs, rew, done, trunc, cmds) if want_cmds else (obs, rew, done, trunc)

    def ssl_env_step(self, task, actions, auto_reset=True, max_steps=1000, want_cmds=False):
        """static defenders / contested possession (5 actions), dribbling (4), pass endurance (3)"""
        a = np.ascontiguousarray(actions, dtype=np.float32).reshape(self.n, TASK_ACT[task])
        obs = np.zeros((self.n, task_obs_dim(task, self.nb, self.ny)))
        rew = np.zeros(self.n)
        done = np.zeros(self.n, dtype=np.uint8)
        trunc = np.zeros(self.n, dtype=np.uint8)
        cmds = np.zeros((self.n, self.R, 8)) if want_cmds else None
        fn = self.L.orc_ssl_hw_env_step if task in (TASK_SSL_DRIBBLING, TASK_SSL_PASS_ENDURANCE) else self.L.orc_ssl_env_step
        fn(self.h, task, _p(a), int(auto_reset), max_steps, _p(obs), _p(rew), _p(done), _p(trunc), _p(cmds))
        return (obs, rew, done, trunc, cmds) if want_cmds else (obs, rew, done, trunc)

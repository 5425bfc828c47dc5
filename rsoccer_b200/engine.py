"""BatchedWorld -- N independent robot-soccer matches on one B200, behind the C ABI.

This is the host-side owner of what the reference keeps inside one ``robosim.VSS`` /
``robosim.SSL`` object (rsoccer_gym/Simulators/rsim.py:116-124, 169-177), for N matches
at once.  torch supplies device memory and streams only; every computation is a
hand-written sm_100a kernel in ``csrc/`` reached through ``include/rsoccer_b200.h``.
"""
import ctypes as C

import torch

from . import _lib

KIND_VSS, KIND_SSL = 0, 1
TASK_VSS_V0 = _lib.TASK_VSS_V0
TASK_SSL_STATIC_DEFENDERS_V0 = _lib.TASK_SSL_STATIC_DEFENDERS_V0
TASK_SSL_CONTESTED_POSSESSION_V0 = _lib.TASK_SSL_CONTESTED_POSSESSION_V0
TASK_SSL_DRIBBLING_V0 = _lib.TASK_SSL_DRIBBLING_V0
TASK_SSL_PASS_ENDURANCE_V0 = _lib.TASK_SSL_PASS_ENDURANCE_V0

FIELD_KEYS = (
    "length", "width", "penalty_length", "penalty_width", "goal_width", "goal_depth",
    "ball_radius", "rbt_distance_center_kicker", "rbt_kicker_thickness", "rbt_kicker_width",
    "rbt_wheel0_angle", "rbt_wheel1_angle", "rbt_wheel2_angle", "rbt_wheel3_angle",
    "rbt_radius", "rbt_wheel_radius", "rbt_motor_max_rpm",
)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class BatchedWorld:
    """N matches of one world kind, state SoA fp32 in HBM (layout: DESIGN.md section 2)."""

    def __init__(self, kind, field_type, n_blue, n_yellow, time_step_ms=25, n_envs=1, device=None,
                 seed=0, env_offset=0):
        if not torch.cuda.is_available():
            raise _lib.RsError("rsoccer_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.L = _lib.lib()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise _lib.RsError("rsoccer_b200 worlds live on a CUDA device, not on %r" % (self.device,))
        if self.device.index is None:          # "cuda" = the current device, resolved once
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.kind, self.field_type = int(kind), int(field_type)
        self.n_blue, self.n_yellow = int(n_blue), int(n_yellow)
        self.R = self.n_blue + self.n_yellow
        self.K = 6 if kind == KIND_VSS else 11
        self.cmd_dim = 2 if kind == KIND_VSS else 8
        self.n = int(n_envs)
        self.h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.L.rs_create(kind, field_type, n_blue, n_yellow, int(time_step_ms), self.n,
                                        self.device.index, int(seed), int(env_offset), C.byref(self.h)),
                       "rs_create")
            nbytes = self.L.rs_state_bytes(self.h)
            # caller-owned state memory: one torch allocation, SoA views below
            self.state = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            _lib.check(self.L.rs_bind_state(self.h, _ptr(self.state), self._stream()), "rs_bind_state")
        offs = (C.c_int64 * _lib.ARR_COUNT)()
        np_ = C.c_int64()
        _lib.check(self.L.rs_layout(self.h, offs, C.byref(np_)), "rs_layout")
        self.np = int(np_.value)
        R, Np = self.R, self.np

        def view(arr, nbytes, dtype, shape):
            o = int(offs[arr])
            return self.state[o:o + nbytes].view(dtype).view(*shape)

        # zero-copy SoA views (x, y, vx, vy) / (theta, omega): "returns torch CUDA tensors"
        self.body = view(_lib.ARR_BODY, 16 * (R + 1) * Np, torch.float32, (R + 1, Np, 4))
        self.ang = view(_lib.ARR_ANG, 8 * R * Np, torch.float32, (R, Np, 2))
        self.ou = view(_lib.ARR_OU, 8 * max(R - 1, 1) * Np, torch.float32, (max(R - 1, 1), Np, 2))
        # one float task word per match: the previous ball potential (VSS-v0), checkpoints_count
        # (SSLDribbling-v0, dribbling.py:56) or stopped_steps (SSLPassEndurance-v0, pass_endurance.py:57)
        self.task_word = view(_lib.ARR_PREV, 4 * Np, torch.float32, (Np,))
        self.prev_pot = self.task_word
        # raw step word: bits 0-23 = episode step count, bit 24 = "previous potential valid" (VSS-v0)
        self.steps_raw = view(_lib.ARR_STEPS, 4 * Np, torch.int32, (Np,))
        self.info = view(_lib.ARR_INFO, 4 * 9 * Np, torch.float32, (9, Np))

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.rs_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _f32(self, x, shape):
        t = torch.as_tensor(x, dtype=torch.float32, device=self.device).contiguous()
        if tuple(t.shape) != tuple(shape):
            t = t.reshape(shape)
        return t

    @property
    def t(self):
        return int(self.L.rs_get_t(self.h))

    @t.setter
    def t(self, v):
        _lib.check(self.L.rs_set_t(self.h, int(v)), "rs_set_t")

    def sync_t(self):
        """Read the device-resident step counter back (needed after CUDA-graph replays)."""
        _lib.check(self.L.rs_sync_t(self.h, self._stream()), "rs_sync_t")
        return self.t

    @property
    def steps(self):
        """episode step count per match (a copy: the flag bit of the raw word is masked off)"""
        return self.steps_raw & 0xFFFFFF

    def set_option(self, option, value):
        """rs_set_option: e.g. set_option(_lib.OPT_STEP_OVERLAP, 1), see include/rsoccer_b200.h"""
        _lib.check(self.L.rs_set_option(self.h, int(option), int(value)), "rs_set_option")

    def get_option(self, option):
        v = C.c_int64()
        _lib.check(self.L.rs_get_option(self.h, int(option), C.byref(v), self._stream()), "rs_get_option")
        return int(v.value)

    def state_written(self):
        """call after writing the state through the zero-copy views while step overlap is on"""
        self.set_option(_lib.OPT_STEP_OVERLAP, self.get_option(_lib.OPT_STEP_OVERLAP))

    @property
    def launches(self):
        return int(self.L.rs_launch_count(self.h))

    @property
    def kernel_flags(self):
        """bit 0 / 1: task / rs_step kernels run one lane per body; bit 2: compile-time physics constants"""
        return int(self.L.rs_kernel_flags(self.h))

    # ------------------------------------------------------------------ robosim surface
    def field_params(self):
        out = (C.c_double * 17)()
        _lib.check(self.L.rs_field_params(self.h, out), "rs_field_params")
        return dict(zip(FIELD_KEYS, list(out)))

    def reset(self, ball, blue, yellow, mask=None):
        """robosim.reset(ball[x,y,vx,vy], blue[x,y,theta_deg], yellow[...]) per env."""
        b = self._f32(ball, (self.n, 4))
        bl = self._f32(blue, (self.n, self.n_blue, 3)) if self.n_blue else None
        ye = self._f32(yellow, (self.n, self.n_yellow, 3)) if self.n_yellow else None
        m = None if mask is None else torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
        _lib.check(self.L.rs_reset(self.h, _ptr(b), _ptr(bl), _ptr(ye), _ptr(m), self._stream()), "rs_reset")

    def step(self, cmds):
        """robosim.step(cmds): [N, R, 2] (VSS wheel rad/s) or [N, R, 8] (SSL)."""
        c = self._f32(cmds, (self.n, self.R, self.cmd_dim))
        _lib.check(self.L.rs_step(self.h, _ptr(c), self._stream()), "rs_step")

    def get_state(self, out=None):
        """robosim.get_state(): [N, 5 + K R] in the Entities/Frame.py layout."""
        if out is None:
            out = torch.empty(self.n, 5 + self.K * self.R, dtype=torch.float32, device=self.device)
        _lib.check(self.L.rs_get_state(self.h, _ptr(out), self._stream()), "rs_get_state")
        return out

    def set_raw(self, raw):
        r = self._f32(raw, (self.n, 4 + 6 * self.R))
        _lib.check(self.L.rs_set_raw(self.h, _ptr(r), self._stream()), "rs_set_raw")

    def get_raw(self):
        out = torch.empty(self.n, 4 + 6 * self.R, dtype=torch.float32, device=self.device)
        _lib.check(self.L.rs_get_raw(self.h, _ptr(out), self._stream()), "rs_get_raw")
        return out

    # ------------------------------------------------------------------ task level
    def obs_dim(self, task):
        d = self.L.rs_task_obs_dim(self.h, task)
        if d < 0:
            _lib.check(d, "rs_task_obs_dim")
        return d

    def alloc_outputs(self, task):
        n, d = self.n, self.obs_dim(task)
        return (torch.empty(n, d, dtype=torch.float32, device=self.device),
                torch.empty(n, dtype=torch.float32, device=self.device),
                torch.empty(n, dtype=torch.uint8, device=self.device),
                torch.empty(n, dtype=torch.uint8, device=self.device))

    def alloc_host_outputs(self, task):
        """Pinned host (obs, reward, done, trunc) for the *_host steps, views of ONE block laid
        out [obs | reward | done | trunc] so that the library moves them in a single copy."""
        n, d = self.n, self.obs_dim(task)
        blk = torch.empty(4 * n * d + 4 * n + 2 * n, dtype=torch.uint8).pin_memory()
        o0, o1, o2 = 4 * n * d, 4 * n * d + 4 * n, 4 * n * d + 5 * n
        return (blk[:o0].view(torch.float32).view(n, d), blk[o0:o1].view(torch.float32),
                blk[o1:o2], blk[o2:])

    def task_reset(self, task, mask=None, obs=None):
        if obs is None:
            obs = torch.zeros(self.n, self.obs_dim(task), dtype=torch.float32, device=self.device)
        m = None if mask is None else torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
        _lib.check(self.L.rs_task_reset(self.h, task, _ptr(m), _ptr(obs), self._stream()), "rs_task_reset")
        return obs

    def vss_env_step(self, actions, normals=None, auto_reset=True, max_steps=1200, out=None,
                     cmds_out=None):
        a = self._f32(actions, (self.n, 2))
        z = None if normals is None else self._f32(normals, (self.n, 2 * (self.R - 1)))
        obs, rew, done, trunc = out if out is not None else self.alloc_outputs(TASK_VSS_V0)
        _lib.check(self.L.rs_vss_env_step(self.h, _ptr(a), _ptr(z), int(auto_reset), int(max_steps),
                                          _ptr(obs), _ptr(rew), _ptr(done), _ptr(trunc),
                                          _ptr(cmds_out), self._stream()), "rs_vss_env_step")
        return obs, rew, done, trunc

    def ssl_env_step(self, task, actions, auto_reset=True, max_steps=1000, out=None, cmds_out=None):
        """static defenders / contested possession (5 actions), dribbling (4), pass endurance (3)"""
        a = self._f32(actions, (self.n, int(self.L.rs_task_act_dim(task))))
        obs, rew, done, trunc = out if out is not None else self.alloc_outputs(task)
        _lib.check(self.L.rs_ssl_env_step(self.h, task, _ptr(a), int(auto_reset), int(max_steps),
                                          _ptr(obs), _ptr(rew), _ptr(done), _ptr(trunc),
                                          _ptr(cmds_out), self._stream()), "rs_ssl_env_step")
        return obs, rew, done, trunc

    # host-buffer end-to-end entry points (numpy / pinned torch CPU tensors in and out)
    def vss_env_step_host(self, h_actions, h_obs, h_rew, h_done, h_trunc, auto_reset=True, max_steps=1200):
        _lib.check(self.L.rs_vss_env_step_host(
            self.h, C.c_void_p(h_actions.data_ptr()), int(auto_reset), int(max_steps),
            C.c_void_p(h_obs.data_ptr()), C.c_void_p(h_rew.data_ptr()), C.c_void_p(h_done.data_ptr()),
            C.c_void_p(h_trunc.data_ptr()), self._stream()), "rs_vss_env_step_host")

    def ssl_env_step_host(self, task, h_actions, h_obs, h_rew, h_done, h_trunc, auto_reset=True,
                          max_steps=1000):
        _lib.check(self.L.rs_ssl_env_step_host(
            self.h, task, C.c_void_p(h_actions.data_ptr()), int(auto_reset), int(max_steps),
            C.c_void_p(h_obs.data_ptr()), C.c_void_p(h_rew.data_ptr()), C.c_void_p(h_done.data_ptr()),
            C.c_void_p(h_trunc.data_ptr()), self._stream()), "rs_ssl_env_step_host")

    # split-phase forms (rs_*_env_step_host_begin / rs_host_step_wait): enqueue on the current stream and
    # return; host_step_wait() blocks until the outputs have landed.  The buffers belong to the library until then.
    def vss_env_step_host_begin(self, h_actions, h_obs, h_rew, h_done, h_trunc, auto_reset=True, max_steps=1200):
        _lib.check(self.L.rs_vss_env_step_host_begin(
            self.h, C.c_void_p(h_actions.data_ptr()), int(auto_reset), int(max_steps),
            C.c_void_p(h_obs.data_ptr()), C.c_void_p(h_rew.data_ptr()), C.c_void_p(h_done.data_ptr()),
            C.c_void_p(h_trunc.data_ptr()), self._stream()), "rs_vss_env_step_host_begin")

    def ssl_env_step_host_begin(self, task, h_actions, h_obs, h_rew, h_done, h_trunc, auto_reset=True,
                                max_steps=1000):
        _lib.check(self.L.rs_ssl_env_step_host_begin(
            self.h, task, C.c_void_p(h_actions.data_ptr()), int(auto_reset), int(max_steps),
            C.c_void_p(h_obs.data_ptr()), C.c_void_p(h_rew.data_ptr()), C.c_void_p(h_done.data_ptr()),
            C.c_void_p(h_trunc.data_ptr()), self._stream()), "rs_ssl_env_step_host_begin")

    def host_step_wait(self):
        _lib.check(self.L.rs_host_step_wait(self.h), "rs_host_step_wait")

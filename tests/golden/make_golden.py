#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the UNMODIFIED reference classes.

    python tests/golden/make_golden.py            (needs /root/reference; CPU only)

What is pinned: the reference's own Python task logic -- command conversion
(vss_gym.py:119-142, 235-254; static_defenders.py:114-148), observation
(vss_gym.py:93-117; static_defenders.py:90-112), reward / done (vss_gym.py:144-192, 256-311;
static_defenders.py:150-212, 256-322; contested_possession.py:136-208), the OU process
(Utils/Utils.py:5-24), the adapter's row packing (Simulators/rsim.py:91-102, 128-155) and
the Frame parser (Entities/Frame.py:17-93) -- executed by importing rsoccer_gym from
/root/reference with import-level stand-ins for gymnasium / pygame and with `robosim`
served by the CPU oracle (tests/golden/shims/robosim.py).  What is NOT pinned: the physics
inside robosim.step (parity unpinned, see oracle/rs_oracle.c header) -- the state rows in
these files are the oracle's own.

Per step the files hold everything needed to replay the step in isolation: raw state and
task state BEFORE the step, the agent action, the standard normals the reference drew,
and the reference's outputs (sent commands, observation, reward, done, state AFTER).
"""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("RS_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "shims"))

import robosim as oracle_robosim  # noqa: E402  (tests/golden/shims)
from rsoccer_b200 import compat  # noqa: E402

compat.install(robosim_module=oracle_robosim, force_shims=True)
sys.path.insert(0, REF)
import gymnasium as gym  # noqa: E402  (the stand-in)
import rsoccer_gym  # noqa: E402,F401  (the reference, unmodified)

DEG = 180.0 / np.pi


def raw_of(env):
    return env.unwrapped.rsim.simulator._w.get_raw()[0].copy()


def set_raw(env, raw):
    env.unwrapped.rsim.simulator._w.set_raw(raw.reshape(1, -1))
    u = env.unwrapped
    u.frame = u.rsim.get_frame()


class NormalLog:
    """records every np.random.normal draw the reference makes (OU noise)."""

    def __init__(self):
        self.orig = np.random.normal
        self.draws = []

    def __enter__(self):
        def normal(*a, **k):
            v = self.orig(*a, **k)
            self.draws.append(np.array(v, dtype=np.float64).reshape(-1))
            return v
        np.random.normal = normal
        return self

    def __exit__(self, *a):
        np.random.normal = self.orig


def run_vss(seed, steps, scripted):
    random.seed(seed)
    np.random.seed(seed)
    rng = np.random.default_rng(seed)
    env = gym.make("VSS-v0")
    u = env.unwrapped
    rec = {k: [] for k in ("raw_before", "ou_before", "prev_pot", "has_prev", "steps_before", "action",
                           "normals", "cmds", "obs", "reward", "done", "trunc", "state_after", "raw_after")}
    obs0, _ = env.reset()
    rec0 = dict(reset_obs=obs0.astype(np.float64), reset_state=np.array(u.rsim.simulator.get_state()))
    for t in range(steps):
        if scripted and t % 40 == 20:      # send the ball towards a goal so that `done` fires
            raw = raw_of(env)
            side = 1.0 if (t // 40) % 2 == 0 else -1.0
            raw[0], raw[1], raw[2], raw[3] = side * 0.66, rng.uniform(-0.1, 0.1), side * 1.5, 0.0
            set_raw(env, raw)
        if scripted and t % 40 == 35:      # one step short of the registered limit (rsoccer_gym/__init__.py:4): TimeLimit fires
            u.steps = 1199
            env._elapsed_steps = 1199
        rec["raw_before"].append(raw_of(env))
        rec["ou_before"].append(np.concatenate([np.asarray(u.ou_actions[i].x_prev, dtype=np.float64) for i in range(1, 6)]))
        rec["prev_pot"].append(0.0 if u.previous_ball_potential is None else float(u.previous_ball_potential))
        rec["has_prev"].append(0 if u.previous_ball_potential is None else 1)
        rec["steps_before"].append(u.steps)
        a = rng.uniform(-1, 1, 2).astype(np.float32)
        if t % 7 == 3:
            a[:] = rng.uniform(-0.06, 0.06, 2)       # exercise the deadzone
        with NormalLog() as log:
            obs, rew, term, trunc, info = env.step(a)
        rec["action"].append(a.astype(np.float64))
        rec["normals"].append(np.concatenate(log.draws))
        cm = np.zeros((6, 2))
        for c in u.sent_commands:
            cm[(3 + c.id) if c.yellow else c.id] = (c.v_wheel0, c.v_wheel1)
        rec["cmds"].append(cm.reshape(-1))
        rec["obs"].append(obs.astype(np.float64))
        rec["reward"].append(float(rew)); rec["done"].append(int(term)); rec["trunc"].append(int(trunc))
        rec["state_after"].append(np.array(u.rsim.simulator.get_state()))
        rec["raw_after"].append(raw_of(env))
        if term or trunc:
            env.reset()
    out = {k: np.array(v) for k, v in rec.items()}
    out.update(rec0)
    out["info_keys"] = np.array(list(info.keys()))
    out["info_last"] = np.array([float(v) for v in info.values()])
    out["field"] = np.array([getattr(u.field, k) for k in u.field.__dataclass_fields__])
    out["max_pos_v_w"] = np.array([u.max_pos, u.max_v, u.max_w])
    return out


def run_ssl(env_id, seed, steps, scripted):
    random.seed(seed)
    np.random.seed(seed)
    rng = np.random.default_rng(seed)
    env = gym.make(env_id)
    u = env.unwrapped
    R = u.n_robots_blue + u.n_robots_yellow
    rec = {k: [] for k in ("raw_before", "steps_before", "action", "cmds", "obs", "reward", "done", "trunc",
                           "state_after", "raw_after")}
    env.reset()
    for t in range(steps):
        if scripted and t % 30 == 10:       # robot right behind the ball, facing the goal
            raw = raw_of(env)
            ang = rng.uniform(-0.4, 0.4)
            raw[4] = raw[0] - 0.2 * np.cos(ang); raw[5] = raw[1] - 0.2 * np.sin(ang); raw[6] = ang
            raw[7:10] = 0.0
            set_raw(env, raw)
        rec["raw_before"].append(raw_of(env))
        rec["steps_before"].append(u.steps)
        a = rng.uniform(-1, 1, 5).astype(np.float32)
        if scripted:
            raw = raw_of(env)
            d = raw[0:2] - raw[4:6]
            a[0:2] = 0.5 * d / (np.linalg.norm(d) + 1e-9)
            a[2] = 0.0
            a[4] = 1.0
            a[3] = 1.0 if t % 30 == 25 else -1.0
        obs, rew, term, trunc, info = env.step(a)
        rec["action"].append(a.astype(np.float64))
        cm = np.zeros((R, 8))
        for c in u.sent_commands:
            row = (u.n_robots_blue + c.id) if c.yellow else c.id
            cm[row] = (c.wheel_speed, c.v_x, c.v_y, c.v_theta, 0.0, c.kick_v_x, c.kick_v_z, c.dribbler)
        rec["cmds"].append(cm.reshape(-1))
        rec["obs"].append(obs.astype(np.float64))
        rec["reward"].append(float(rew)); rec["done"].append(int(term)); rec["trunc"].append(int(trunc))
        rec["state_after"].append(np.array(u.rsim.simulator.get_state()))
        rec["raw_after"].append(raw_of(env))
        if term or trunc:
            env.reset()
    out = {k: np.array(v) for k, v in rec.items()}
    out["field"] = np.array([getattr(u.field, k) for k in u.field.__dataclass_fields__])
    out["info_keys"] = np.array(list(info.keys()))
    return out


BRANCH_KEYS = ("goal", "rbt_in_gk_area", "done_ball_out", "done_ball_out_right", "done_rbt_out", "collision")


def run_ssl_branches(env_id, seed, reps=4):
    """Every branch of the done chain (static_defenders.py:179-198, contested_possession.py:165-191) and a
    TimeLimit truncation, `reps` times each: one scripted frame per recorded step.  `branch` = index in
    BRANCH_KEYS of the reward_shaping_total counter the reference incremented (-1: none)."""
    random.seed(seed)
    np.random.seed(seed)
    rng = np.random.default_rng(seed)
    env = gym.make(env_id)
    u = env.unwrapped
    nb, ny = u.n_robots_blue, u.n_robots_yellow
    R = nb + ny
    limit = env._max_episode_steps
    scen = ["goal", "ball_out_right", "ball_out_left", "ball_out_side", "rbt_out_left", "rbt_out_side",
            "rbt_in_gk", "trunc", "shaping"] + (["collision"] if "Contested" in env_id else [])
    rec = {k: [] for k in ("raw_before", "steps_before", "action", "cmds", "obs", "reward", "done", "trunc",
                           "state_after", "raw_after", "branch")}
    names = []
    for rep in range(reps):
        for name in scen:
            env.reset()
            raw = raw_of(env)
            j = lambda s=0.02: rng.uniform(-s, s)                       # noqa: E731
            for k in range(ny):                                        # defenders parked along the lower touch line
                raw[4 + 6 * (nb + k):4 + 6 * (nb + k) + 6] = (0.4 + 0.4 * k, -1.7, 0.0, 0.0, 0.0, 0.0)
            rb = [1.0 + j(0.2), 0.5 + j(0.2), rng.uniform(-3, 3), 0.0, 0.0, 0.0]
            ball = [1.5 + j(0.2), 0.2 + j(0.2), 0.0, 0.0]
            side = 1.0 if rep % 2 == 0 else -1.0
            if name == "goal":
                ball = [2.97 + j(0.01), 0.25 * side + j(), 3.0, j(0.2)]
            elif name == "ball_out_right":
                ball = [2.97 + j(0.01), 1.0 * side + j(0.3), 3.0, j(0.2)]
            elif name == "ball_out_left":
                ball = [0.03 + j(0.01), 0.3 * side + j(0.2), -3.0, j(0.2)]
            elif name == "ball_out_side":
                ball = [1.0 + j(0.3), 1.97 * side, j(0.2), 3.0 * side]
            elif name == "rbt_out_left":
                rb[0], rb[1] = -0.25 + j(0.02), 0.3 * side
            elif name == "rbt_out_side":
                rb[0], rb[1] = 1.0 + j(0.3), 2.05 * side
            elif name == "rbt_in_gk":
                rb[0], rb[1] = 2.5 + j(0.1), 0.4 * side + j(0.2)
            elif name == "collision":
                yx, yy = 1.5 + j(0.2), 0.6 * side
                raw[4 + 6 * nb:4 + 6 * nb + 6] = (yx, yy, np.pi, 0.0, 0.0, 0.0)
                rb = [yx - 0.185, yy + j(0.01), 0.0, 1.5, 0.0, 0.0]
                ball = [0.8, -0.8 * side, 0.0, 0.0]
            raw[0:4] = ball
            raw[4:10] = rb
            set_raw(env, raw)
            if name == "trunc":
                u.steps = limit - 1
                env._elapsed_steps = limit - 1
            rec["raw_before"].append(raw_of(env))
            rec["steps_before"].append(u.steps)
            a = rng.uniform(-1, 1, 5).astype(np.float32)
            before = dict(u.reward_shaping_total or {})
            obs, rew, term, trunc, info = env.step(a)
            inc = [i for i, k in enumerate(BRANCH_KEYS) if info.get(k, 0) - before.get(k, 0) > 0]
            assert len(inc) <= 2, (name, inc)
            rec["branch"].append(inc[0] if inc else -1)
            rec["action"].append(a.astype(np.float64))
            cm = np.zeros((R, 8))
            for c in u.sent_commands:
                row = (nb + c.id) if c.yellow else c.id
                cm[row] = (c.wheel_speed, c.v_x, c.v_y, c.v_theta, 0.0, c.kick_v_x, c.kick_v_z, c.dribbler)
            rec["cmds"].append(cm.reshape(-1))
            rec["obs"].append(obs.astype(np.float64))
            rec["reward"].append(float(rew)); rec["done"].append(int(term)); rec["trunc"].append(int(trunc))
            rec["state_after"].append(np.array(u.rsim.simulator.get_state()))
            rec["raw_after"].append(raw_of(env))
            names.append(name)
    out = {k: np.array(v) for k, v in rec.items()}
    out["scenario"] = np.array(names)
    out["branch_keys"] = np.array(BRANCH_KEYS)
    out["field"] = np.array([getattr(u.field, k) for k in u.field.__dataclass_fields__])
    return out


def run_ssl_hw(env_id, seed, steps, scripted):
    """SSLDribbling-v0 / SSLPassEndurance-v0 (dribbling.py, pass_endurance.py): same record as
    run_ssl plus the per-episode counter of the task before / after the step (dribbling:
    checkpoints_count; pass endurance: stopped_steps) and pass endurance's reward_shaping_total."""
    random.seed(seed)
    np.random.seed(seed)
    rng = np.random.default_rng(seed)
    env = gym.make(env_id)
    u = env.unwrapped
    drib = env_id == "SSLDribbling-v0"
    R = u.n_robots_blue + u.n_robots_yellow
    n_act = 4 if drib else 3
    rec = {k: [] for k in ("raw_before", "steps_before", "counter_before", "action", "cmds", "obs", "reward",
                           "done", "trunc", "state_after", "raw_after", "counter_after", "info_after")}
    env.reset()

    def counter():
        return u.checkpoints_count if drib else u.stopped_steps

    for t in range(steps):
        a = rng.uniform(-1, 1, n_act).astype(np.float32)
        if scripted and drib and t % 6 == 3:
            # the ball about to cross y = 0 inside the x window of checkpoint cc (or just outside it),
            # in the rewarded or in the reversed direction; the robot parked away from it
            cc = int(rng.integers(0, 7))
            lo, hi = {0: (-1.0, -0.5), 1: (-1.5, -1.0)}.get(cc, (-2.0, -1.5) if cc % 2 == 0 else (-3.0, -2.0))
            raw = raw_of(env)
            down = (cc % 2 == 0) != (rng.uniform() < 0.25)
            raw[0] = rng.uniform(lo - 0.1, hi + 0.1); raw[1] = 0.004 if down else -0.004
            raw[2] = 0.0; raw[3] = -2.0 if down else 2.0
            raw[4] = raw[0] + 0.4; raw[5] = 0.5; raw[7:10] = 0.0
            if rng.uniform() < 0.15:
                raw[4] = 1.02                      # out of the course
            set_raw(env, raw)
            u.checkpoints_count = cc
            a[:] = 0.0
        if scripted and not drib and t % 8 == 4:
            # the ball rolling into the receiver's mouth (infrared -> +1, done), or drifting out of the box
            raw = raw_of(env)
            rx, ry, rth = raw[10], raw[11], raw[12]
            d = rng.uniform(0.13, 0.3)
            off = rng.uniform(-0.02, 0.02) if rng.uniform() < 0.7 else rng.uniform(0.3, 0.6)
            raw[0] = rx + d * np.cos(rth) - off * np.sin(rth); raw[1] = ry + d * np.sin(rth) + off * np.cos(rth)
            raw[2] = -2.5 * np.cos(rth); raw[3] = -2.5 * np.sin(rth)
            set_raw(env, raw)
        if scripted and not drib and t % 8 != 4:
            a[0] = rng.uniform(-0.2, 0.2); a[2] = 1.0
            a[1] = 1.0 if t % 8 == 1 else rng.uniform(-0.6, 0.6)
        rec["raw_before"].append(raw_of(env))
        rec["steps_before"].append(u.steps)
        rec["counter_before"].append(counter())
        obs, rew, term, trunc, info = env.step(a.copy())        # pass_endurance.py:102 edits the action in place
        rec["action"].append(a.astype(np.float64))
        cm = np.zeros((R, 8))
        for c in u.sent_commands:
            row = (u.n_robots_blue + c.id) if c.yellow else c.id
            cm[row] = (c.wheel_speed, c.v_x, c.v_y, c.v_theta, 0.0, c.kick_v_x, c.kick_v_z, c.dribbler)
        rec["cmds"].append(cm.reshape(-1))
        rec["obs"].append(obs.astype(np.float64))
        rec["reward"].append(float(rew)); rec["done"].append(int(term)); rec["trunc"].append(int(trunc))
        rec["state_after"].append(np.array(u.rsim.simulator.get_state()))
        rec["raw_after"].append(raw_of(env))
        rec["counter_after"].append(counter())
        rst = getattr(u, "reward_shaping_total", None) or {}
        rec["info_after"].append([float(rst.get("reversed_dist", 0.0)), float(rst.get("ball_grad", 0.0))])
        if term or trunc:
            env.reset()
    out = {k: np.array(v) for k, v in rec.items()}
    out["field"] = np.array([getattr(u.field, k) for k in u.field.__dataclass_fields__])
    return out


def main():
    np.savez_compressed(os.path.join(HERE, "vss_v0_random.npz"), **run_vss(11, 260, False))
    np.savez_compressed(os.path.join(HERE, "vss_v0_goals.npz"), **run_vss(12, 200, True))
    np.savez_compressed(os.path.join(HERE, "ssl_static_defenders_random.npz"),
                        **run_ssl("SSLStaticDefenders-v0", 21, 200, False))
    np.savez_compressed(os.path.join(HERE, "ssl_static_defenders_fetch.npz"),
                        **run_ssl("SSLStaticDefenders-v0", 22, 240, True))
    np.savez_compressed(os.path.join(HERE, "ssl_contested_possession_random.npz"),
                        **run_ssl("SSLContestedPossession-v0", 31, 200, False))
    np.savez_compressed(os.path.join(HERE, "ssl_contested_possession_fetch.npz"),
                        **run_ssl("SSLContestedPossession-v0", 32, 240, True))
    np.savez_compressed(os.path.join(HERE, "ssl_static_defenders_branches.npz"),
                        **run_ssl_branches("SSLStaticDefenders-v0", 23))
    np.savez_compressed(os.path.join(HERE, "ssl_contested_possession_branches.npz"),
                        **run_ssl_branches("SSLContestedPossession-v0", 33))
    np.savez_compressed(os.path.join(HERE, "ssl_dribbling_random.npz"), **run_ssl_hw("SSLDribbling-v0", 41, 200, False))
    np.savez_compressed(os.path.join(HERE, "ssl_dribbling_course.npz"), **run_ssl_hw("SSLDribbling-v0", 42, 300, True))
    np.savez_compressed(os.path.join(HERE, "ssl_pass_endurance_random.npz"),
                        **run_ssl_hw("SSLPassEndurance-v0", 51, 200, False))
    np.savez_compressed(os.path.join(HERE, "ssl_pass_endurance_catch.npz"),
                        **run_ssl_hw("SSLPassEndurance-v0", 52, 300, True))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            d = np.load(os.path.join(HERE, f))
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes; steps", len(d["reward"]),
                  "dones", int(d["done"].sum()), "truncs", int(d["trunc"].sum()))


if __name__ == "__main__":
    main()

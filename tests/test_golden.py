"""Golden fixtures produced by the UNMODIFIED reference task code (tests/golden/make_golden.py)
replayed step by step through (a) the CPU oracle and (b) the CUDA path (-m gpu).

Pins commands / observation / reward / done / wire state against the reference's own
Python (vss_gym.py, static_defenders.py, contested_possession.py, rsim.py, Frame.py)."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
G = os.path.join(HERE, "golden")

VSS_FILES = ["vss_v0_random.npz", "vss_v0_goals.npz"]
SSL_FILES = [("ssl_static_defenders_random.npz", 1, 1, 6, 1000), ("ssl_static_defenders_fetch.npz", 1, 1, 6, 1000),
             ("ssl_static_defenders_branches.npz", 1, 1, 6, 1000),
             ("ssl_contested_possession_random.npz", 2, 1, 1, 1200), ("ssl_contested_possession_fetch.npz", 2, 1, 1, 1200),
             ("ssl_contested_possession_branches.npz", 2, 1, 1, 1200)]
# reward_shaping_total counters of the done chain, in the order make_golden.BRANCH_KEYS records them, and the
# column of each in the engine's info block (rs_spec.h RS_SSL_INFO order)
BRANCH_INFO_COL = {"goal": 0, "rbt_in_gk_area": 1, "done_ball_out": 2, "done_ball_out_right": 3, "done_rbt_out": 4,
                   "collision": 8}
# SSLDribbling-v0 (task 3: 1 blue + 4 yellow) and SSLPassEndurance-v0 (task 4: 2 blue)
HW_FILES = [("ssl_dribbling_random.npz", 3, 1, 4, 4800), ("ssl_dribbling_course.npz", 3, 1, 4, 4800),
            ("ssl_pass_endurance_random.npz", 4, 2, 0, 1200), ("ssl_pass_endurance_catch.npz", 4, 2, 0, 1200)]


def _load(name):
    return np.load(os.path.join(G, name))


def test_fixtures_fire_every_branch_of_the_reference_done_chains():
    """static_defenders.py:179-198 and contested_possession.py:165-191: robot out, robot in the goalkeeper
    area, ball out, ball out over the goal line, goal (+5), collision, plus the TimeLimit truncation of
    rsoccer_gym/__init__.py -- each recorded at least 3 times from the unmodified reference classes; VSS-v0:
    both goals (+-10, vss_gym.py:161-170) and the 1200-step truncation."""
    for name, extra in (("ssl_static_defenders_branches.npz", ()), ("ssl_contested_possession_branches.npz", ("collision",))):
        d = _load(name)
        keys = [str(k) for k in d["branch_keys"]]
        for k in ("goal", "rbt_in_gk_area", "done_ball_out", "done_ball_out_right", "done_rbt_out") + extra:
            assert (d["branch"] == keys.index(k)).sum() >= 3, (name, k)
        assert d["trunc"].sum() >= 3 and (d["reward"] == 5).sum() >= 3, name
        assert ((d["done"] == 0) & (d["trunc"] == 0)).sum() >= 3, name            # and the shaping branch
        # a truncated step that is not also terminated
        assert ((d["trunc"] == 1) & (d["done"] == 0)).sum() >= 3, name
    d = _load("vss_v0_goals.npz")
    assert (d["reward"] == 10).sum() >= 2 and (d["reward"] == -10).sum() >= 2 and d["trunc"].sum() >= 3


def test_field_params_match_reference_field(oracle):
    """Field(**get_field_params()) round trip recorded from the reference (rsim.py:49-50)."""
    d = _load("vss_v0_random.npz")
    w = oracle.OracleWorld(0, 0, 3, 3)
    assert np.allclose(d["field"], list(w.field_params().values()))
    # base-env derived normalisers, vss_gym_base.py:52-58
    mp, mv, mw = d["max_pos_v_w"]
    assert abs(mp - 0.9) < 1e-12 and abs(mv - 440 / 60 * 2 * np.pi * 0.026) < 1e-12
    assert abs(mw - np.rad2deg(mv / 0.04)) < 1e-9


@pytest.mark.parametrize("name", VSS_FILES)
def test_oracle_vss_env_step_vs_reference(oracle, name):
    d = _load(name)
    T = len(d["reward"])
    w = oracle.OracleWorld(0, 0, 3, 3, 25, T)       # one env per recorded step, all replayed at once
    w.set_raw(d["raw_before"])
    w.set_task_state(ou=d["ou_before"], prev_pot=d["prev_pot"], has_prev=d["has_prev"],
                     steps=d["steps_before"], info=np.zeros((T, 9)))
    obs, rew, done, trunc, cmds = w.vss_env_step(d["action"].astype(np.float32), normals=d["normals"],
                                                 auto_reset=False, max_steps=1200, want_cmds=True)
    assert np.abs(cmds.reshape(T, -1) - d["cmds"]).max() < 2e-5        # rad/s; reference path is fp32 for the agent
    assert (done == d["done"]).all() and (trunc == d["trunc"]).all()
    assert np.abs(w.get_state() - d["state_after"]).max() < 1e-3      # 1e-3: deg/s columns after fp32-vs-fp64 commands
    assert np.abs(w.get_raw() - d["raw_after"]).max() < 5e-6         # the reference converts the agent action in fp32
    assert np.abs(obs - d["obs"]).max() < 2e-6                         # reference casts obs to float32
    assert np.abs(rew - d["reward"]).max() < 1e-6
    assert d["done"].sum() >= (4 if "goals" in name else 0)


@pytest.mark.parametrize("name,task,nb,ny,max_steps", SSL_FILES)
def test_oracle_ssl_env_step_vs_reference(oracle, name, task, nb, ny, max_steps):
    d = _load(name)
    T = len(d["reward"])
    R = nb + ny
    w = oracle.OracleWorld(1, 2, nb, ny, 25, T)
    w.set_raw(d["raw_before"])
    w.set_task_state(steps=d["steps_before"], info=np.zeros((T, 9)))
    obs, rew, done, trunc, cmds = w.ssl_env_step(task, d["action"].astype(np.float32), auto_reset=False,
                                                 max_steps=max_steps, want_cmds=True)
    assert np.abs(cmds.reshape(T, -1) - d["cmds"]).max() < 1e-6
    assert (done == d["done"]).all() and (trunc == d["trunc"]).all()
    assert np.abs(w.get_state() - d["state_after"]).max() < 1e-4      # deg/s and wheel rad/s columns; fp32 action path in the reference
    assert np.abs(obs - d["obs"]).max() < 2e-6
    assert np.abs(rew - d["reward"]).max() < 1e-6
    if "branch" in d.files:          # the oracle takes the same branch of the done chain as the reference did
        info, keys = w.get_task_state()["info"], [str(k) for k in d["branch_keys"]]
        for i, k in enumerate(keys):
            rows = d["branch"] == i
            assert (info[rows, BRANCH_INFO_COL[k]] >= 1).all(), (name, k)
        quiet = d["branch"] == -1
        assert (info[quiet][:, [0, 1, 2, 3, 4, 8]] == 0).all()
    if "fetch" in name:
        assert d["done"].sum() >= 3 and (d["obs"][:, 11] > 0.5).sum() > 5     # infrared seen


@pytest.mark.parametrize("name,task,nb,ny,max_steps", HW_FILES)
def test_oracle_ssl_hw_env_step_vs_reference(oracle, name, task, nb, ny, max_steps):
    """dribbling.py / pass_endurance.py: commands, observation (taken before the reward updates
    the checkpoint counter, ssl_gym_base.py:83-85), reward, done and the task counter."""
    d = _load(name)
    T = len(d["reward"])
    w = oracle.OracleWorld(1, 2, nb, ny, 25, T)
    w.set_raw(d["raw_before"])
    w.set_task_state(steps=np.maximum(d["steps_before"], 1), prev_pot=d["counter_before"].astype(np.float64),
                     info=np.zeros((T, 9)))
    obs, rew, done, trunc, cmds = w.ssl_env_step(task, d["action"].astype(np.float32), auto_reset=False,
                                                 max_steps=max_steps, want_cmds=True)
    assert np.abs(cmds.reshape(T, -1) - d["cmds"]).max() < 1e-6
    assert (done == d["done"]).all() and (trunc == d["trunc"]).all()
    assert np.abs(w.get_state() - d["state_after"]).max() < 1e-4
    assert np.abs(obs - d["obs"]).max() < 2e-6
    assert np.abs(rew - d["reward"]).max() < 1e-6
    ts = w.get_task_state()
    assert (ts["prev_pot"] == d["counter_after"]).all()
    if task == 4:
        dn = d["done"] == 1
        assert np.abs(ts["info"][dn, 0] - d["info_after"][dn, 0]).max() < 1e-9       # reversed_dist
    if "course" in name:
        assert d["reward"].sum() >= 10 and d["done"].sum() >= 5 and len(set(d["counter_after"])) >= 6
    if "catch" in name:
        assert (d["reward"] > 0.9).sum() >= 5 and (d["reward"] < -0.5).sum() >= 5


@pytest.mark.gpu
@pytest.mark.parametrize("name", VSS_FILES)
def test_cuda_vss_env_step_vs_reference(engine, name):
    import torch
    d = _load(name)
    T = len(d["reward"])
    g = engine.BatchedWorld(0, 0, 3, 3, 25, T)
    g.set_raw(d["raw_before"].astype(np.float32))
    g.ou[:, :T, :] = torch.tensor(d["ou_before"].reshape(T, 5, 2), dtype=torch.float32).permute(1, 0, 2).cuda()
    g.task_word[:T] = torch.tensor(d["prev_pot"], dtype=torch.float32).cuda()
    g.steps_raw[:T] = torch.tensor(d["steps_before"] | (d["has_prev"] << 24), dtype=torch.int32).cuda()
    cg = torch.zeros(T, 6, 2, device="cuda")
    obs, rew, done, trunc = g.vss_env_step(d["action"].astype(np.float32), normals=d["normals"].astype(np.float32),
                                           auto_reset=False, max_steps=1200, cmds_out=cg)
    assert np.abs(cg.cpu().numpy().reshape(T, -1) - d["cmds"]).max() < 1e-4
    assert (done.cpu().numpy() == d["done"]).all() and (trunc.cpu().numpy() == d["trunc"]).all()
    assert np.abs(obs.cpu().numpy() - d["obs"]).max() < 1e-4
    assert np.abs(rew.cpu().numpy() - d["reward"]).max() < 2e-4
    st = g.get_state().cpu().numpy()
    err = np.abs(st - d["state_after"])
    for r in range(6):
        err[:, 5 + 6 * r + 2] = np.abs((st[:, 5 + 6 * r + 2] - d["state_after"][:, 5 + 6 * r + 2] + 180) % 360 - 180)
        err[:, 5 + 6 * r + 2] /= 100.0     # degrees: 1e-4 rad = 5.7e-3 deg
        err[:, 5 + 6 * r + 5] /= 100.0
    assert err.max() < 1e-4, err.max()


def _ssl_obs_tol(nobs, nb, vel_scale, dribbling=False):
    """per-column tolerance of an SSL observation row against the recorded reference row: 1e-4 on positions and
    sin / cos (the reference normalises positions by max_pos > 1), 1e-4 x vel_scale on velocities (DESIGN.md
    section 5: fp32 contact normals on a 6 m field), and on v_theta the same scaled by 57.3 / 10 -- the reference
    divides deg/s by max_w = 10 -- the infrared flag exact."""
    tol = np.full(nobs, 1e-4)
    o = 1 if dribbling else 0
    tol[o + 2:o + 4] = 1e-4 * vel_scale
    for r in range(nb):
        b = o + 4 + 8 * r
        tol[b + 4:b + 6] = 1e-4 * vel_scale
        tol[b + 6] = 1e-4 * vel_scale * 57.3 / 10.0
        tol[b + 7] = 0.0
    return tol


def _oracle_margin(oracle, d, task, nb, ny, max_steps, hw):
    """the recorded steps replayed in the fp64 oracle: which rows does IT call ill conditioned (a decision --
    contact, kicker box, a done / checkpoint line -- taken by < 5e-6 m)?  Only those may differ."""
    T = len(d["reward"])
    w = oracle.OracleWorld(1, 2, nb, ny, 25, T)
    w.set_raw(d["raw_before"].astype(np.float32).astype(np.float64))
    if hw:
        w.set_task_state(steps=np.maximum(d["steps_before"], 1), prev_pot=d["counter_before"].astype(np.float64),
                         info=np.zeros((T, 9)))
    else:
        w.set_task_state(steps=d["steps_before"], info=np.zeros((T, 9)))
    w.ssl_env_step(task, d["action"].astype(np.float32), auto_reset=False, max_steps=max_steps)
    return w.margin()


@pytest.mark.gpu
@pytest.mark.parametrize("name,task,nb,ny,max_steps", SSL_FILES)
def test_cuda_ssl_env_step_vs_reference(engine, oracle, name, task, nb, ny, max_steps):
    """every recorded reference step through the fused CUDA step: commands, done, truncation, observation and
    reward within the documented tolerance on EVERY row the oracle does not flag -- no flips, no percentiles"""
    import torch
    d = _load(name)
    T = len(d["reward"])
    R = nb + ny
    g = engine.BatchedWorld(1, 2, nb, ny, 25, T)
    g.set_raw(d["raw_before"].astype(np.float32))
    g.steps_raw[:T] = torch.tensor(d["steps_before"], dtype=torch.int32).cuda()
    cg = torch.zeros(T, R, 8, device="cuda")
    obs, rew, done, trunc = g.ssl_env_step(task, d["action"].astype(np.float32), auto_reset=False,
                                           max_steps=max_steps, cmds_out=cg)
    assert np.abs(cg.cpu().numpy().reshape(T, -1) - d["cmds"]).max() < 1e-4
    ok = _oracle_margin(oracle, d, task, nb, ny, max_steps, False) >= 5e-6
    vs = max(1.0, float(d["field"][0]) / 2)
    print("%s: %d of %d rows flagged by the oracle" % (name, int((~ok).sum()), T))
    assert (~ok).mean() <= 0.03
    assert (done.cpu().numpy()[ok] == d["done"][ok]).all() and (trunc.cpu().numpy() == d["trunc"]).all()
    eo = np.abs(obs.cpu().numpy() - d["obs"])
    assert (eo[ok] <= _ssl_obs_tol(eo.shape[1], nb, vs)).all(), np.argwhere(eo[ok] > _ssl_obs_tol(eo.shape[1], nb, vs))[:5]
    assert np.abs(rew.cpu().numpy() - d["reward"])[ok].max() <= 1e-4 * vs


@pytest.mark.gpu
@pytest.mark.parametrize("name,task,nb,ny,max_steps", HW_FILES)
def test_cuda_ssl_hw_env_step_vs_reference(engine, oracle, name, task, nb, ny, max_steps):
    """SSLDribbling-v0 / SSLPassEndurance-v0 recorded from the unmodified reference classes,
    replayed through the fused CUDA step (k_ssl_hw_env_step); same bar as above."""
    import torch
    d = _load(name)
    T = len(d["reward"])
    R = nb + ny
    g = engine.BatchedWorld(1, 2, nb, ny, 25, T)
    g.set_raw(d["raw_before"].astype(np.float32))
    g.steps_raw[:T] = torch.tensor(np.maximum(d["steps_before"], 1), dtype=torch.int32).cuda()
    g.task_word[:T] = torch.tensor(d["counter_before"], dtype=torch.float32).cuda()
    cg = torch.zeros(T, R, 8, device="cuda")
    obs, rew, done, trunc = g.ssl_env_step(task, d["action"].astype(np.float32), auto_reset=False,
                                           max_steps=max_steps, cmds_out=cg)
    assert np.abs(cg.cpu().numpy().reshape(T, -1) - d["cmds"]).max() < 1e-4
    ok = _oracle_margin(oracle, d, task, nb, ny, max_steps, True) >= 5e-6
    vs = max(1.0, float(d["field"][0]) / 2)
    print("%s: %d of %d rows flagged by the oracle" % (name, int((~ok).sum()), T))
    assert (~ok).mean() <= 0.12       # the dribbling course starts with the ball resting on the mouth plane
    assert (done.cpu().numpy()[ok] == d["done"][ok]).all() and (trunc.cpu().numpy() == d["trunc"]).all()
    eo = np.abs(obs.cpu().numpy() - d["obs"])
    if task == 3:
        tol = _ssl_obs_tol(eo.shape[1], nb, vs, dribbling=True)
        tol[12] = 0.0                                     # infrared as +-1
    else:                                                 # pass endurance rows: x y sin cos v_theta infrared per robot
        tol = np.full(eo.shape[1], 1e-4)
        tol[2:4] = 1e-4 * vs
        for r in range(nb):
            tol[4 + 6 * r + 4] = 1e-4 * vs * 57.3 / 10.0
            tol[4 + 6 * r + 5] = 0.0
    assert (eo[ok] <= tol).all(), np.argwhere(eo[ok] > tol)[:5]
    assert np.abs(rew.cpu().numpy() - d["reward"])[ok].max() <= 1e-4 * vs
    assert (g.task_word[:T].cpu().numpy()[ok] == d["counter_after"][ok]).all()

"""-m gpu: the CUDA path (through the C ABI) against the CPU oracle, same seeded inputs."""
import numpy as np
import pytest
import torch

from parity import (RESYNC_CASES, RESYNC_EPS, RESYNC_MAX_FLAGGED, TOL, check_close, random_cmds, random_raw, raw_diff,
                    resynced_scene)

pytestmark = pytest.mark.gpu


def _worlds(E, O, kind, ft, nb, ny, n, seed=7, env_offset=0):
    g = E.BatchedWorld(kind, ft, nb, ny, 25, n, seed=seed, env_offset=env_offset)
    o = O.OracleWorld(kind, ft, nb, ny, 25, n, seed=seed, env_offset=env_offset, threads=8)
    return g, o


@pytest.mark.parametrize("kind,ft,nb,ny", RESYNC_CASES)
def test_step_parity_resynced(engine, oracle, kind, ft, nb, ny):
    """rs_step vs oracle, one control step from identical random (contact-rich) states.  Every env the
    oracle does not flag (a decision taken by < 2e-5 m that triggers an impulse) must agree to 1e-4; the
    flagged fraction is bounded by its measured value + 2 points (tests/parity.py RESYNC_MAX_FLAGGED)."""
    E, O = engine, oracle
    n, R = (4096 if nb + ny <= 10 else 1000), nb + ny      # 11 v 11: the largest world the reference can build
    g, o = _worlds(E, O, kind, ft, nb, ny, n)
    fp = o.field_params()
    rng = np.random.default_rng(1234 + 10 * kind + R)
    worst, worst_fl = 0.0, 0.0
    for it in range(6):
        raw, cmds = resynced_scene(rng, kind, n, R, fp, it)
        g.set_raw(raw); o.set_raw(raw)
        g.step(cmds); o.step(cmds.astype(np.float64))
        vs = max(1.0, fp["length"] / 2) if kind == 1 else 1.0
        err = raw_diff(g.get_raw().cpu().numpy(), o.get_raw(), R, vel_scale=vs)
        w, fl = check_close(err, o.margin(), "step kind=%d R=%d it=%d" % (kind, R, it),
                            max_flagged=RESYNC_MAX_FLAGGED[(kind, ft, nb, ny)], eps=RESYNC_EPS)
        worst, worst_fl = max(worst, w), max(worst_fl, fl)
        # the wire format agrees too (degrees, infrared, wheel speeds)
        sg, so = g.get_state().cpu().numpy().astype(np.float64), o.get_state()
        K = 6 if kind == 0 else 11
        ok = o.margin() >= RESYNC_EPS
        ds = np.abs(sg - so)
        for r in range(R):
            c = 5 + K * r + 2
            ds[:, c] = np.abs((sg[:, c] - so[:, c] + 180.0) % 360.0 - 180.0)
        tol = np.full(sg.shape[1], TOL * vs)
        for r in range(R):
            tol[5 + K * r + 2] = 1e-2          # degrees
            tol[5 + K * r + 5] = 2e-2          # degrees / s
            if kind == 1:
                tol[5 + K * r + 7:5 + K * r + 11] = 5e-3 * vs   # wheel rad/s = v / 0.02475
        assert (ds[ok] <= tol).all(), "get_state mismatch %s" % np.argwhere(ds[ok] > tol)[:5]
    print("worst abs err %.3e, flagged fraction %.4f" % (worst, worst_fl))


def test_step_parity_free_running(engine, oracle):
    """40 control steps (one simulated second) without re-sync: envs in which no decision came within 1 mm
    (91 % of them: robots curving up the field at 0.1-0.3 m/s, a rolling ball) stay within 1e-3."""
    E, O = engine, oracle
    n, R = 2048, 6
    g, o = _worlds(E, O, 0, 0, 3, 3, n)
    rng = np.random.default_rng(5)
    ball = np.zeros((n, 4)); ball[:, 0] = rng.uniform(-0.1, 0.1, n); ball[:, 1] = 0.55
    ball[:, 2] = rng.uniform(-0.3, 0.3, n)
    xs = np.array([-0.5, -0.3, -0.1, 0.1, 0.3, 0.5])
    rob = np.zeros((n, 6, 3)); rob[:, :, 0] = xs; rob[:, :, 1] = rng.uniform(-0.3, 0.2, (n, 6))
    rob[:, :, 2] = rng.uniform(60, 120, (n, 6))
    ball = ball.astype(np.float32); rob = rob.astype(np.float32)
    g.reset(ball, rob[:, :3], rob[:, 3:]); o.reset(ball, rob[:, :3], rob[:, 3:])
    cmds = rng.uniform(4, 12, (n, 6, 2)).astype(np.float32)
    cmds[:, :, 1] = cmds[:, :, 0] + rng.uniform(-1, 1, (n, 6))
    mmin = np.full(n, 1e9)
    for _ in range(40):
        g.step(cmds); o.step(cmds.astype(np.float64))
        mmin = np.minimum(mmin, o.margin())
    err = raw_diff(g.get_raw().cpu().numpy(), o.get_raw(), R)
    ok = mmin > 1e-3
    assert ok.mean() > 0.85
    print("free running, 40 steps: max abs err %.3e over %.1f %% of the envs" % (err[ok].max(), 100 * ok.mean()))
    assert err[ok].max() < 1e-3, err[ok].max()


def _sync_task(g, o, R):
    raw = g.get_raw().cpu().numpy().astype(np.float64)
    o.set_raw(raw)
    st = g.steps_raw[:g.n].cpu().numpy()
    ou = g.ou[:, :g.n, :].permute(1, 0, 2).reshape(g.n, -1).cpu().numpy().astype(np.float64)
    o.set_task_state(ou=ou[:, :2 * (R - 1)], prev_pot=g.prev_pot[:g.n].cpu().numpy().astype(np.float64),
                     has_prev=((st >> 24) & 1).astype(np.int32), steps=(st & 0xFFFFFF).astype(np.int32),
                     info=g.info[:, :g.n].t().cpu().numpy().astype(np.float64))
    o.t = g.t


def test_vss_env_step_parity(engine, oracle):
    """Fused VSSEnv.step (Philox OU noise, reward, done, auto-reset, obs) vs the oracle,
    re-synced each step, 60 steps, with short episodes so auto-reset is exercised."""
    E, O = engine, oracle
    n, R = 4096, 6
    g, o = _worlds(E, O, 0, 0, 3, 3, n, seed=99, env_offset=1000)
    g.task_reset(E.TASK_VSS_V0)
    _sync_task(g, o, R)
    # device placement == oracle placement on the same stream
    o2 = O.OracleWorld(0, 0, 3, 3, 25, n, seed=99, env_offset=1000)
    o2.task_reset(O.TASK_VSS)
    assert raw_diff(g.get_raw().cpu().numpy(), o2.get_raw(), R).max() < 1e-5
    rng = np.random.default_rng(3)
    worst = {"obs": 0.0, "rew": 0.0, "raw": 0.0}
    n_done = 0
    for it in range(60):
        if it == 20:   # speed things up: balls heading for the goals
            raw = g.get_raw().cpu().numpy()
            raw[:, 0] = rng.uniform(0.55, 0.7, n) * rng.choice([-1, 1], n)
            raw[:, 1] = rng.uniform(-0.15, 0.15, n)
            raw[:, 2] = np.sign(raw[:, 0]) * rng.uniform(0.5, 2.0, n)
            g.set_raw(raw)
        _sync_task(g, o, R)
        act = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
        cg = torch.zeros(n, R, 2, device="cuda")
        obs, rew, done, trunc = g.vss_env_step(act, max_steps=25, cmds_out=cg)
        oobs, orew, odone, otrunc, ocmd = o.vss_env_step(act, max_steps=25, want_cmds=True)
        m = o.margin()
        ok = m >= 5e-6
        assert (np.abs(cg.cpu().numpy() - ocmd)[ok].max()) < 2e-3       # rad/s
        assert (done.cpu().numpy()[ok] == odone[ok]).all()
        assert (trunc.cpu().numpy() == otrunc).all()
        n_done += int(odone.sum() + otrunc.sum())
        e_obs = np.abs(obs.cpu().numpy() - oobs).max(axis=1)
        e_rew = np.abs(rew.cpu().numpy() - orew)
        e_raw = raw_diff(g.get_raw().cpu().numpy(), o.get_raw(), R)
        worst["obs"] = max(worst["obs"], check_close(e_obs, m, "obs it=%d" % it)[0])
        worst["rew"] = max(worst["rew"], check_close(e_rew, m, "reward it=%d" % it, tol=2e-4)[0])
        worst["raw"] = max(worst["raw"], check_close(e_raw, m, "state it=%d" % it)[0])
        ts = o.get_task_state()
        st = g.steps_raw[:n].cpu().numpy()
        assert ((st & 0xFFFFFF)[ok] == ts["steps"][ok]).all()
        gi = g.info[:, :n].t().cpu().numpy()
        assert np.abs(gi - ts["info"])[ok].max() < 2e-3
    assert n_done > n          # every env was reset at least once on average
    print("worst", worst, "resets", n_done)


@pytest.mark.parametrize("task,nb,ny,max_steps", [(1, 1, 6, 30), (2, 1, 1, 30)])
def test_ssl_env_step_parity(engine, oracle, task, nb, ny, max_steps):
    E, O = engine, oracle
    n, R = 4096, nb + ny
    g, o = _worlds(E, O, 1, 2, nb, ny, n, seed=5, env_offset=77)
    g.task_reset(task)
    o2 = O.OracleWorld(1, 2, nb, ny, 25, n, seed=5, env_offset=77)
    o2.task_reset(task)
    assert raw_diff(g.get_raw().cpu().numpy(), o2.get_raw(), R).max() < 1e-5
    rng = np.random.default_rng(11)
    worst = {"obs": 0.0, "rew": 0.0, "raw": 0.0}
    n_done = n_goal_or_out = n_infra = 0
    grp = rng.integers(0, 3, n)     # 0: random, 1: flee over x < -0.2, 2: fetch the ball and shoot
    for it in range(50):
        if it % 10 == 5:      # group 2: robot 0.25 m behind the ball, facing it
            raw = g.get_raw().cpu().numpy()
            k = grp == 2
            ang = rng.uniform(-0.5, 0.5, n)
            raw[k, 4] = (raw[:, 0] - 0.25 * np.cos(ang))[k]
            raw[k, 5] = (raw[:, 1] - 0.25 * np.sin(ang))[k]
            raw[k, 6] = ang[k]
            raw[k, 7:10] = 0.0
            g.set_raw(raw)
        _sync_task(g, o, R)
        act = rng.uniform(-1, 1, (n, 5)).astype(np.float32)
        act[grp == 1, 0] = -1.0
        raw = g.get_raw().cpu().numpy()
        d = raw[:, 0:2] - raw[:, 4:6]
        k = grp == 2
        act[k, 0:2] = (d / (np.linalg.norm(d, axis=1, keepdims=True) + 1e-6))[k] * 0.6
        act[k, 2] = 0.0
        act[k, 4] = 1.0
        act[k, 3] = 1.0 if it % 10 == 9 else -1.0
        cg = torch.zeros(n, R, 8, device="cuda")
        obs, rew, done, trunc = g.ssl_env_step(task, act, max_steps=max_steps, cmds_out=cg)
        oobs, orew, odone, otrunc, ocmd = o.ssl_env_step(task, act, max_steps=max_steps, want_cmds=True)
        m = o.margin()
        ok = m >= 5e-6
        assert np.abs(cg.cpu().numpy() - ocmd)[ok].max() < 1e-4
        assert (done.cpu().numpy()[ok] == odone[ok]).all()
        assert (trunc.cpu().numpy() == otrunc).all()
        n_done += int(odone.sum() + otrunc.sum())
        n_goal_or_out += int(odone.sum())
        n_infra += int((oobs[:, 11] > 0.5).sum())
        e_obs = np.abs(obs.cpu().numpy() - oobs).max(axis=1)
        e_rew = np.abs(rew.cpu().numpy() - orew)
        e_raw = raw_diff(g.get_raw().cpu().numpy(), o.get_raw(), R, vel_scale=3.0)
        worst["obs"] = max(worst["obs"], check_close(e_obs, m, "obs it=%d" % it, max_flagged=0.05)[0])
        worst["rew"] = max(worst["rew"], check_close(e_rew, m, "reward it=%d" % it, max_flagged=0.05)[0])
        worst["raw"] = max(worst["raw"], check_close(e_raw, m, "state it=%d" % it, max_flagged=0.05)[0])
    assert n_done > n
    assert n_infra > 0 and n_goal_or_out > n // 4, (n_infra, n_goal_or_out)
    print("worst", worst, "resets", n_done, "infrared steps", n_infra)


@pytest.mark.parametrize("task,nb,ny,max_steps", [(3, 1, 4, 40), (4, 2, 0, 40)])
def test_ssl_hw_env_step_parity(engine, oracle, task, nb, ny, max_steps):
    """Fused SSLDribbling-v0 / SSLPassEndurance-v0 step vs the oracle, re-synced each step, with
    scripted scenes (checkpoint crossings, passes into the receiver's mouth) and short episodes so
    that rewards, every done branch and the on-device auto-reset placement are exercised."""
    E, O = engine, oracle
    n, R = 4096, nb + ny
    nact = 4 if task == 3 else 3
    g, o = _worlds(E, O, 1, 2, nb, ny, n, seed=6, env_offset=33)
    g.task_reset(task)
    o2 = O.OracleWorld(1, 2, nb, ny, 25, n, seed=6, env_offset=33)
    o2.task_reset(task)
    assert raw_diff(g.get_raw().cpu().numpy(), o2.get_raw(), R).max() < 1e-5
    assert np.abs(g.task_reset(task).cpu().numpy() - o2.task_obs(task)).max() < 1e-5
    rng = np.random.default_rng(12)
    worst = {"obs": 0.0, "rew": 0.0, "raw": 0.0}
    n_done = n_rew = 0
    for it in range(50):
        raw = g.get_raw().cpu().numpy()
        k = rng.random(n) < 0.25
        if task == 3:
            # the ball about to cross y = 0 inside (or just outside) the x window of checkpoint cc
            cc = rng.integers(0, 7, n)
            lo = np.where(cc == 0, -1.0, np.where(cc == 1, -1.5, np.where(cc % 2 == 0, -2.0, -3.0)))
            hi = np.where(cc == 0, -0.5, np.where(cc == 1, -1.0, np.where(cc % 2 == 0, -1.5, -2.0)))
            down = (cc % 2 == 0) != (rng.random(n) < 0.25)
            bx = rng.uniform(lo - 0.1, hi + 0.1)
            for node in (-0.5, -1.0, -1.5, -2.0):          # not inside a parked robot
                near = np.abs(bx - node) < 0.13
                bx[near] = node + np.where(bx[near] < node, -0.13, 0.13)
            raw[k, 0] = bx[k]
            raw[k, 1] = np.where(down, 0.004, -0.004)[k]
            raw[k, 2] = 0.0
            raw[k, 3] = np.where(down, -2.0, 2.0)[k]
            raw[k, 4] = (raw[:, 0] + 0.4)[k]; raw[k, 5] = 0.5; raw[k, 7:10] = 0.0
            out = k & (rng.random(n) < 0.1)
            raw[out, 4] = 1.02
            g.set_raw(raw)
            pp = g.prev_pot[:n].cpu().numpy()
            pp[k] = cc[k]
            g.prev_pot[:n] = torch.tensor(pp, device="cuda")
            st = g.steps_raw[:n].cpu().numpy()
            g.steps_raw[:n] = torch.tensor(np.maximum(st, 1), device="cuda")
        else:
            rx, ry, rth = raw[:, 10], raw[:, 11], raw[:, 12]
            dd = rng.uniform(0.13, 0.3, n)
            off = np.where(rng.random(n) < 0.7, rng.uniform(-0.02, 0.02, n), rng.uniform(0.3, 0.6, n))
            raw[k, 0] = (rx + dd * np.cos(rth) - off * np.sin(rth))[k]
            raw[k, 1] = (ry + dd * np.sin(rth) + off * np.cos(rth))[k]
            raw[k, 2] = (-2.5 * np.cos(rth))[k]; raw[k, 3] = (-2.5 * np.sin(rth))[k]
            g.set_raw(raw)
        _sync_task(g, o, R)
        act = rng.uniform(-1, 1, (n, nact)).astype(np.float32)
        if task == 3:
            act[k] = 0.0
        cg = torch.zeros(n, R, 8, device="cuda")
        obs, rew, done, trunc = g.ssl_env_step(task, act, max_steps=max_steps, cmds_out=cg)
        oobs, orew, odone, otrunc, ocmd = o.ssl_env_step(task, act, max_steps=max_steps, want_cmds=True)
        m = o.margin()
        ok = m >= 5e-6
        assert np.abs(cg.cpu().numpy() - ocmd)[ok].max() < 1e-4
        assert (done.cpu().numpy()[ok] == odone[ok]).all()
        assert (trunc.cpu().numpy() == otrunc).all()
        n_done += int(odone.sum() + otrunc.sum())
        n_rew += int((orew > 0.9).sum())
        e_obs = np.abs(obs.cpu().numpy() - oobs).max(axis=1)
        e_rew = np.abs(rew.cpu().numpy() - orew)
        e_raw = raw_diff(g.get_raw().cpu().numpy(), o.get_raw(), R, vel_scale=3.0)
        e_cnt = np.abs(g.prev_pot[:n].cpu().numpy() - o.get_task_state()["prev_pot"])
        worst["obs"] = max(worst["obs"], check_close(e_obs, m, "obs it=%d" % it, max_flagged=0.05)[0])
        worst["rew"] = max(worst["rew"], check_close(e_rew, m, "reward it=%d" % it, max_flagged=0.05)[0])
        worst["raw"] = max(worst["raw"], check_close(e_raw, m, "state it=%d" % it, max_flagged=0.05)[0])
        check_close(e_cnt, m, "task counter it=%d" % it, max_flagged=0.05)
        if task == 4:
            e_inf = np.abs(g.info[:2, :n].t().cpu().numpy() - o.get_task_state()["info"][:, :2]).max(axis=1)
            check_close(e_inf, m, "reward_shaping_total it=%d" % it, tol=2e-4, max_flagged=0.05)
    assert n_done > n and n_rew > n // 8, (n_done, n_rew)
    print("worst", worst, "resets", n_done, "rewards", n_rew)


def test_batch_vs_single_and_shard_invariance(engine):
    """env i of a batch == the same env alone; results do not depend on the sharding."""
    E = engine
    n = 512
    big = E.BatchedWorld(0, 0, 3, 3, 25, n, seed=3, env_offset=0)
    lo = E.BatchedWorld(0, 0, 3, 3, 25, n // 2, seed=3, env_offset=0)
    hi = E.BatchedWorld(0, 0, 3, 3, 25, n // 2, seed=3, env_offset=n // 2)
    for w in (big, lo, hi):
        w.task_reset(E.TASK_VSS_V0)
    g = torch.Generator(device="cpu").manual_seed(0)
    for _ in range(30):
        a = (torch.rand(n, 2, generator=g) * 2 - 1).cuda()
        ob, rb, db, tb = big.vss_env_step(a, max_steps=10)
        ol, rl, dl, tl = lo.vss_env_step(a[:n // 2], max_steps=10)
        oh, rh, dh, th = hi.vss_env_step(a[n // 2:], max_steps=10)
        assert torch.equal(ob, torch.cat([ol, oh]))
        assert torch.equal(rb, torch.cat([rl, rh]))
        assert torch.equal(db, torch.cat([dl, dh]))
    assert torch.equal(big.get_raw(), torch.cat([lo.get_raw(), hi.get_raw()]))


@pytest.mark.parametrize("n", [1, 33, 130])
def test_tiny_and_ragged_worlds_all_tasks(engine, oracle, n):
    """one match, one warp plus one lane, two ragged CTAs: every fused task step vs the oracle
    (both kernel mappings through the `engine` fixture), 12 steps with 5-step episodes."""
    E, O = engine, oracle
    rng = np.random.default_rng(n)
    for task, kind, ft, nb, ny, nact in ((0, 0, 0, 3, 3, 2), (1, 1, 2, 1, 6, 5), (2, 1, 2, 1, 1, 5), (3, 1, 2, 1, 4, 4), (4, 1, 2, 2, 0, 3)):
        R = nb + ny
        g, o = _worlds(E, O, kind, ft, nb, ny, n, seed=21, env_offset=5)
        g.task_reset(task)
        for it in range(12):
            _sync_task(g, o, R)
            act = rng.uniform(-1, 1, (n, nact)).astype(np.float32)
            if task == 0:
                obs, rew, done, trunc = g.vss_env_step(act, max_steps=5)
                oobs, orew, odone, otrunc = o.vss_env_step(act, max_steps=5)
            else:
                obs, rew, done, trunc = g.ssl_env_step(task, act, max_steps=5)
                oobs, orew, odone, otrunc = o.ssl_env_step(task, act, max_steps=5)
            ok = o.margin() >= 5e-6
            assert obs.shape == oobs.shape and (trunc.cpu().numpy() == otrunc).all()
            assert (done.cpu().numpy()[ok] == odone[ok]).all()
            assert np.abs(obs.cpu().numpy() - oobs)[ok].max(initial=0.0) < 2e-4
            assert np.abs(rew.cpu().numpy() - orew)[ok].max(initial=0.0) < 2e-4


@pytest.mark.parametrize("n", [65536, 262144])
def test_full_size_properties(engine, oracle, n):
    """BASELINE.json's full sizes (65 536 matches per GPU; 262 144 = config 5's total), through
    size-independent properties: the whole world == its two halves stepped as separate worlds
    (bit-exact, auto-reset included), every body stays inside the walls, observations inside the
    normalisation bounds, and a random sample of 512 matches of the big world replayed alone in
    the fp64 oracle agrees to the parity tolerance."""
    E, O = engine, oracle
    big = E.BatchedWorld(0, 0, 3, 3, 25, n, seed=13)
    lo = E.BatchedWorld(0, 0, 3, 3, 25, n // 2, seed=13, env_offset=0)
    hi = E.BatchedWorld(0, 0, 3, 3, 25, n // 2, seed=13, env_offset=n // 2)
    for w in (big, lo, hi):
        w.task_reset(E.TASK_VSS_V0)
    g = torch.Generator(device="cpu").manual_seed(2)
    rng = np.random.default_rng(4)
    pick = np.sort(rng.choice(n, 512, replace=False))
    o = O.OracleWorld(0, 0, 3, 3, 25, 512, seed=13, threads=8)
    fp = o.field_params()
    for it in range(24):
        a = (torch.rand(n, 2, generator=g) * 2 - 1).cuda()
        if it % 6 == 5:
            # sampled oracle replay of this step: state, task words and noise of the picked matches
            raw = big.get_raw()[pick].cpu().numpy().astype(np.float64)
            st = big.steps_raw[:n][pick].cpu().numpy()
            ou = big.ou[:, :n, :][:, pick].permute(1, 0, 2).reshape(512, -1).cpu().numpy().astype(np.float64)
            o.set_raw(raw)
            o.set_task_state(ou=ou, prev_pot=big.prev_pot[:n][pick].cpu().numpy().astype(np.float64),
                             has_prev=((st >> 24) & 1).astype(np.int32), steps=(st & 0xFFFFFF).astype(np.int32),
                             info=np.zeros((512, 9)))
            ou_before = big.ou[:, :n, :].clone()
        ob, rb, db, tb = big.vss_env_step(a, max_steps=11)
        ol, rl, dl, tl = lo.vss_env_step(a[:n // 2], max_steps=11)
        oh, rh, dh, th = hi.vss_env_step(a[n // 2:], max_steps=11)
        assert torch.equal(ob, torch.cat([ol, oh])) and torch.equal(rb, torch.cat([rl, rh]))
        assert torch.equal(db, torch.cat([dl, dh])) and torch.equal(tb, torch.cat([tl, th]))
        assert bool(torch.isfinite(ob).all()) and float(ob.abs().max()) <= 1.2 + 1e-6
        if it % 6 == 5:
            # the OU noise the device drew = (new - (1 - theta dt) old) / (sigma sqrt(dt)): feed it to the oracle
            z = (big.ou[:, :n, :] - ou_before * (1 - 0.17 * 0.025)) / (0.5 * np.sqrt(0.025))
            keep = (~(db.bool() | tb.bool()))[pick].cpu().numpy()          # reset matches restart their OU state
            zs = z[:, pick].permute(1, 0, 2).reshape(512, -1).cpu().numpy().astype(np.float64)
            oobs, orew, odone, otr = o.vss_env_step(a[pick].cpu().numpy(), normals=zs, auto_reset=False, max_steps=11)
            ok = keep & (o.margin() >= 5e-6)
            assert ok.mean() > 0.8
            assert np.abs(ob[pick].cpu().numpy() - oobs)[ok].max() < 2e-4
            assert np.abs(rb[pick].cpu().numpy() - orew)[ok].max() < 2e-4
            assert (db[pick].cpu().numpy()[ok] == odone[ok]).all()
    raw = big.get_raw()
    assert torch.equal(raw, torch.cat([lo.get_raw(), hi.get_raw()]))
    x = raw[:, [0] + [4 + 6 * r for r in range(6)]].abs()
    y = raw[:, [1] + [5 + 6 * r for r in range(6)]].abs()
    assert float(x.max()) <= fp["length"] / 2 + fp["goal_depth"] + 1e-6 and float(y.max()) <= fp["width"] / 2 + 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("n", [4096, 40000])
def test_vss_step_keeps_a_world_and_its_mirror_image_mirror_images(engine, n):
    """Size-independent property of the physics itself (rs_step = robosim.VSS.step): the model is
    symmetric under the reflection y -> -y (theta -> -theta, omega -> -omega, left and right wheel
    swapped) and so is its fp32 arithmetic -- negation is exact, |.|, min / max, rsqrt, sin and cos are
    even or odd, the order of pairs and walls does not change -- so a world and its mirror image
    stay mirror images BIT FOR BIT through drive, contacts, walls and goal recesses, in every kernel
    family (40 000 matches: the packed / scalar lane-per-match kernels; lane per body for both)."""
    E = engine
    R = 6
    rng = np.random.default_rng(11)
    a, b = E.BatchedWorld(0, 0, 3, 3, 25, n, seed=1), E.BatchedWorld(0, 0, 3, 3, 25, n, seed=1)
    raw = random_raw(rng, n, R, 0.75, 0.65).astype(np.float32)
    sign = np.ones(4 + 6 * R, dtype=np.float32)
    sign[[1, 3]] = -1.0
    for r in range(R):
        sign[[4 + 6 * r + 1, 4 + 6 * r + 2, 4 + 6 * r + 4, 4 + 6 * r + 5]] = -1.0      # y, theta, vy, omega
    a.set_raw(raw)
    b.set_raw(raw * sign)
    tsign = torch.tensor(sign, device="cuda")
    for _ in range(40):
        c = torch.tensor(random_cmds(rng, 0, n, R), device="cuda")
        a.step(c)
        b.step(c.flip(-1).contiguous())
    ra, rb = a.get_raw(), b.get_raw() * tsign
    bad = int((ra != rb).any(dim=1).sum())
    moved = float((ra[:, :2] - torch.tensor(raw[:, :2], device="cuda")).abs().max())
    assert moved > 0.05                      # the worlds did evolve
    assert bad == 0, "%d of %d matches lost the mirror symmetry" % (bad, n)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [4096, 40000])
def test_vss_env_step_keeps_a_world_and_its_mirror_image_mirror_images(engine, n):
    """The reflection property through the whole fused VSSEnv.step: with the action components swapped and the OU
    normals of each robot swapped (supplied explicitly; no auto-reset, whose placement draws are not mirrored),
    commands, physics, reward, done and the observation rows of a world and of its mirror image stay mirror
    images bit for bit for 30 steps (observation: y, sin(theta), v_y, omega change sign)."""
    E = engine
    R = 6
    g = torch.Generator(device="cpu").manual_seed(5)
    a, b = E.BatchedWorld(0, 0, 3, 3, 25, n, seed=2), E.BatchedWorld(0, 0, 3, 3, 25, n, seed=2)
    a.task_reset(E.TASK_VSS_V0); b.task_reset(E.TASK_VSS_V0)
    sign = torch.ones(4 + 6 * R)
    sign[[1, 3]] = -1.0
    for r in range(R):
        sign[[4 + 6 * r + 1, 4 + 6 * r + 2, 4 + 6 * r + 4, 4 + 6 * r + 5]] = -1.0
    sign = sign.cuda()
    osign = torch.tensor([1, -1, 1, -1] + [1, -1, -1, 1, 1, -1, -1] * 3 + [1, -1, 1, -1, -1] * 3, dtype=torch.float32).cuda()
    b.set_raw(a.get_raw() * sign)
    worst_rew = 0.0
    for _ in range(30):
        act = (torch.rand(n, 2, generator=g) * 2 - 1).cuda()
        z = torch.randn(n, R - 1, 2, generator=g).cuda()
        oa, ra, da, _ = a.vss_env_step(act, normals=z.reshape(n, -1), auto_reset=False, max_steps=10 ** 6)
        ob, rb, db, _ = b.vss_env_step(act.flip(-1).contiguous(), normals=z.flip(-1).reshape(n, -1).contiguous(),
                                       auto_reset=False, max_steps=10 ** 6)
        assert torch.equal(oa, ob * osign) and torch.equal(da, db)
        assert torch.equal(ra, rb)
    assert torch.equal(a.get_raw(), b.get_raw() * sign)

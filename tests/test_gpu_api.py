"""-m gpu: the host-side mirrors of the reference interface on top of the CUDA engine."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_robosim_shim_matches_oracle_shim(engine, oracle):
    """the same call sequence rsim.py makes (ctor, get_field_params, reset, step, get_state) through the
    CUDA-backed `robosim` module and through the oracle-backed one"""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden", "shims"))
    import robosim as O_sim
    from rsoccer_b200.compat import robosim as G_sim
    rng = np.random.default_rng(0)
    for cls, nb, ny, C, ft in (("VSS", 3, 3, 2, 0), ("SSL", 1, 6, 8, 2)):
        R = nb + ny
        ball = [0, 0, 0, 0]
        blue = [[-0.2 * i, 0, 0] for i in range(1, nb + 1)]          # rsim.py:19-24
        yel = [[0.2 * i, 0, 0] for i in range(1, ny + 1)]
        g = getattr(G_sim, cls)(ft, nb, ny, 25, ball, blue, yel)
        o = getattr(O_sim, cls)(ft, nb, ny, 25, ball, blue, yel)
        assert g.get_field_params().keys() == o.get_field_params().keys()
        assert np.allclose(list(g.get_field_params().values()), list(o.get_field_params().values()))
        b = np.array([0.3, 0.1, -0.5, 0.2])
        bl = np.column_stack([rng.uniform(-0.6, -0.1, nb), rng.uniform(-0.5, 0.5, nb), rng.uniform(0, 360, nb)])
        ye = np.column_stack([rng.uniform(0.1, 0.6, ny), np.linspace(-0.5, 0.5, ny), rng.uniform(0, 360, ny)])
        g.reset(b, bl, ye); o.reset(b, bl, ye)
        for _ in range(20):
            cmd = np.zeros((R, C))
            if C == 2:
                cmd[:] = rng.uniform(-30, 30, (R, 2))
            else:
                cmd[:, 1:4] = rng.uniform(-1, 1, (R, 3))
            g.step(cmd); o.step(cmd)
        sg, so = np.asarray(g.get_state()), np.asarray(o.get_state())
        assert sg.dtype == np.float64 and sg.shape == so.shape == (5 + (6 if C == 2 else 11) * R,)
        assert np.abs(sg - so).max() < 5e-3       # degrees / wheel rad/s columns dominate
        K = 6 if C == 2 else 11
        xy = [0, 1, 3, 4] + [5 + K * r + c for r in range(R) for c in (0, 1, 3, 4)]
        assert np.abs(sg[xy] - so[xy]).max() < 1e-4
        with pytest.raises(IndexError):
            g.step(np.zeros((R + 1, C)))


def test_batched_rsim_adapter_reset_send_get(engine):
    """RSimVSS mirror: reset(frame) / send_commands(List[Robot]) / get_frame() with per-env tensors"""
    from rsoccer_b200.entities import Ball, Frame, Robot
    from rsoccer_b200.simulators import RSimVSS
    n = 16
    sim = RSimVSS(0, 3, 3, 25, n_envs=n)
    assert sim.field.length == 1.5 and sim.field.rbt_radius == 0.0375
    fr = Frame()
    fr.ball = Ball(x=torch.linspace(-0.3, 0.3, n), y=0.0)
    for i in range(3):
        fr.robots_blue[i] = Robot(x=-0.5, y=-0.3 + 0.3 * i, theta=torch.full((n,), 90.0))
        fr.robots_yellow[i] = Robot(x=0.5, y=-0.3 + 0.3 * i, theta=180.0)
    sim.reset(fr)
    f0 = sim.get_frame()
    assert torch.allclose(f0.ball.x.cpu(), torch.linspace(-0.3, 0.3, n), atol=1e-6)
    assert torch.allclose(f0.robots_blue[1].theta.cpu(), torch.full((n,), 90.0), atol=1e-3)
    w = torch.linspace(5, 36, n)
    for _ in range(10):
        sim.send_commands([Robot(yellow=True, id=2, v_wheel0=w, v_wheel1=w)])
    f1 = sim.get_frame()
    assert torch.allclose(f1.robots_yellow[2].v_x.cpu(), -w * 0.026, atol=1e-4)    # facing 180 deg
    assert f1.robots_blue[0].v_x.abs().max() < 1e-6                                # K12 zero rows
    with pytest.raises(IndexError):
        sim.send_commands([Robot(yellow=False, id=3, v_wheel0=1.0, v_wheel1=1.0)])
    sim.stop()


def test_user_subclass_of_the_batched_base_env(engine):
    """the reference README's example env (README.md:71-113), written against the batched base class"""
    from rsoccer_b200.entities import Ball, Frame, Robot
    from rsoccer_b200.envs import SSLBaseVecEnv

    class Example(SSLBaseVecEnv):
        def __init__(self, num_envs):
            super().__init__(field_type=0, n_robots_blue=1, n_robots_yellow=0, time_step=0.025, num_envs=num_envs)

        def _frame_to_observations(self):
            b, r = self.frame.ball, self.frame.robots_blue[0]
            return torch.stack([b.x, b.y, r.x, r.y], dim=1)

        def _get_commands(self, actions):
            return [Robot(yellow=False, id=0, v_x=actions[:, 0], v_y=actions[:, 1])]

        def _calculate_reward_and_done(self):
            goal = (self.frame.ball.x > self.field.length / 2) & (self.frame.ball.y.abs() < self.field.goal_width / 2)
            return goal.float(), goal

        def _get_initial_positions_frame(self):
            f = Frame()
            f.ball = Ball(x=self.field.length / 2 - self.field.penalty_length, y=0.0)
            f.robots_blue[0] = Robot(x=0.0, y=0.0, theta=0.0)
            return f

    env = Example(8)
    obs, _ = env.reset()
    assert obs.shape == (8, 4) and torch.allclose(obs[:, 0].cpu(), torch.full((8,), 3.5))
    a = torch.zeros(8, 2, device=obs.device); a[:, 0] = torch.linspace(0.5, 2.0, 8)
    for _ in range(40):
        obs, rew, done, trunc, _ = env.step(a)
    assert (obs[1:, 2] > obs[:-1, 2]).all() and obs[:, 2].min() > 0.3
    assert not done.any() and rew.sum() == 0
    env.close()


def test_fused_vec_env_api_and_host_path(engine):
    from rsoccer_b200 import envs
    n = 1000                        # not a multiple of 32 or 64: ragged last warp / CTA
    env = envs.make("VSS-v0", num_envs=n, seed=4)
    ref = envs.make("VSS-v0", num_envs=n, seed=4)
    obs, info = env.reset()
    ref.reset()
    assert obs.shape == (n, 40) and obs.dtype == torch.float32 and obs.is_cuda
    assert set(info) == {"goal_score", "move", "ball_grad", "energy", "goals_blue", "goals_yellow"}
    g = torch.Generator().manual_seed(0)
    for t in range(30):
        a = torch.rand(n, 2, generator=g) * 2 - 1
        o1, r1, d1, t1, i1 = env.step(a.cuda())
        o2, r2, d2, t2 = ref.step_host(a.numpy())                  # numpy in / numpy out, pinned staging
        assert o1.abs().max() <= 1.2 + 1e-6 and d1.dtype == torch.bool
        assert np.array_equal(o1.cpu().numpy(), o2) and np.array_equal(r1.cpu().numpy(), r2)
        assert np.array_equal(d1.cpu().numpy(), d2) and np.array_equal(t1.cpu().numpy(), t2)
    assert float(i1["energy"].max()) < 0 and i1["move"].shape == (n,)
    f = env.frame
    assert f.ball.x.shape == (n,) and len(f.robots_blue) == 3
    with pytest.raises(KeyError):
        envs.make("SSLGoToBall-v0")
    for eid, od, ad, key in (("SSLStaticDefenders-v0", 24, 5, "collision"), ("SSLContestedPossession-v0", 14, 5, "collision"),
                             ("SSLDribbling-v0", 21, 4, None), ("SSLPassEndurance-v0", 16, 3, "reversed_dist")):
        e = envs.make(eid, num_envs=70)
        o, _ = e.reset()
        assert o.shape == (70, od)
        o, r, d, tr, i = e.step(torch.zeros(70, ad, device="cuda"))
        assert o.shape == (70, od) and (key is None or key in i)
        o2, r2, d2, t2 = e.step_host(np.zeros((70, ad), dtype=np.float32))
        assert o2.shape == (70, od)
        e.close()


def test_time_limit_truncation_and_autoreset(engine):
    from rsoccer_b200 import envs
    env = envs.make("VSS-v0", num_envs=64, max_episode_steps=5)
    env.reset()
    a = torch.zeros(64, 2, device="cuda")
    flags = [bool(env.step(a)[3].all()) for _ in range(12)]
    assert flags == [False] * 4 + [True] + [False] * 4 + [True] + [False] * 2
    assert int(env.world.steps[:64].max()) == 2 and int(env.world.steps_raw[:64].max()) >> 24 == 1
    env2 = envs.make("VSS-v0", num_envs=64, max_episode_steps=5, auto_reset=False)
    env2.reset()
    tr = [bool(env2.step(a)[3].all()) for _ in range(7)]
    assert tr == [False] * 4 + [True] * 3


def test_cuda_graph_replay_draws_fresh_noise(engine):
    """the Philox step counter lives in device memory: replays of one captured launch differ"""
    E = engine
    n = 256
    w = E.BatchedWorld(0, 0, 3, 3, 25, n, seed=1)
    w.task_reset(E.TASK_VSS_V0)
    a = torch.zeros(n, 2, device="cuda")
    out = w.alloc_outputs(E.TASK_VSS_V0)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            w.vss_env_step(a, out=out)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            w.vss_env_step(a, out=out)
        ou = []
        for _ in range(4):
            g.replay()
            s.synchronize()
            ou.append(w.ou[:, :n].clone())
    # 3 eager steps + 4 replays ran on the device (the capture itself executes nothing; the
    # host mirror counted it, which is why rs_sync_t exists)
    assert w.t == 4 and w.sync_t() == 7
    inc = [(ou[i + 1] - ou[i] * (1 - 0.17 * 0.025)) for i in range(3)]
    assert not torch.allclose(inc[0], inc[1]) and not torch.allclose(inc[1], inc[2])
    # and the replayed steps equal eager steps of a twin world
    w2 = E.BatchedWorld(0, 0, 3, 3, 25, n, seed=1)
    w2.task_reset(E.TASK_VSS_V0)
    for _ in range(7):
        w2.vss_env_step(a)
    assert torch.equal(w2.get_raw(), w.get_raw())


def test_checkpoint_restore_via_state_tensor(engine):
    """the state is one torch tensor -> clone/restore is a full checkpoint"""
    E = engine
    w = E.BatchedWorld(0, 0, 3, 3, 25, 128, seed=9)
    w.task_reset(E.TASK_VSS_V0)
    a = torch.rand(128, 2, device="cuda") * 2 - 1
    for _ in range(5):
        w.vss_env_step(a)
    snap, t = w.state.clone(), w.t
    ref = [w.vss_env_step(a)[0].clone() for _ in range(5)]
    w.state.copy_(snap); w.t = t
    again = [w.vss_env_step(a)[0].clone() for _ in range(5)]
    assert all(torch.equal(x, y) for x, y in zip(ref, again))


def test_preset_kernel_is_selected_and_matches_the_generic_one(engine, monkeypatch):
    """VSS field 0 at 25 ms steps on kernels whose physics constants are compile-time
    immediates (rs_kernel_flags bit 2).  Same constants bit for bit (checked by the library
    before it selects them), but the compiler contracts multiply-adds differently around
    immediates, so against the run-time-constant kernels (RS_NO_PRESET=1) the results agree to
    rounding, not to the bit: one step from identical state, 1e-5, auto-reset included; a
    1-ulp flip of a contact decision may move a handful of rows further."""
    E = engine
    monkeypatch.setenv("RS_PER_MATCH", "1")
    n = 1000
    a = torch.rand(n, 2, device="cuda") * 2 - 1
    w = E.BatchedWorld(0, 0, 3, 3, 25, n, seed=3)
    assert w.kernel_flags & 4 and not (w.kernel_flags & 1)
    monkeypatch.setenv("RS_NO_PRESET", "1")
    g = E.BatchedWorld(0, 0, 3, 3, 25, n, seed=3)
    assert not (g.kernel_flags & 4)
    monkeypatch.delenv("RS_NO_PRESET")
    assert not (E.BatchedWorld(0, 0, 3, 3, 20, 64, seed=3).kernel_flags & 4)
    w.task_reset(E.TASK_VSS_V0)
    bad = 0
    for _ in range(40):
        g.state.copy_(w.state); g.t = w.t
        ow = w.vss_env_step(a, max_steps=15)
        og = g.vss_env_step(a, max_steps=15)
        row_ok = ((ow[0] - og[0]).abs().amax(dim=1) < 1e-5) & ((ow[1] - og[1]).abs() < 1e-5) & (ow[2] == og[2])
        assert torch.equal(ow[3], og[3])
        bad += int((~row_ok).sum())
    assert bad <= 4, bad


def test_packed_and_scalar_instruction_forms_agree_bit_for_bit(engine, monkeypatch):
    """Large VSS-v0 worlds run the lane-per-match kernel built on the packed fp32x2 forms
    (FFMA2 / FADD2 / FMUL2, rs_kernel_flags bit 3), small ones the scalar forms (shorter dependent
    chains).  Each packed half rounds like the scalar instruction and both kernels are compiled from
    one float2 source, so the choice -- made by world size, hence different for a world and for
    its shards -- must not change a single bit: 60 steps with contacts, goals and auto-reset."""
    E = engine
    monkeypatch.setenv("RS_PER_MATCH", "1")
    n = 3000
    a = torch.rand(n, 2, device="cuda") * 2 - 1
    monkeypatch.setenv("RS_PACKED", "1")
    p = E.BatchedWorld(0, 0, 3, 3, 25, n, seed=5)
    assert (p.kernel_flags & 12) == 12
    monkeypatch.setenv("RS_PACKED", "0")
    q = E.BatchedWorld(0, 0, 3, 3, 25, n, seed=5)
    assert (q.kernel_flags & 12) == 4
    monkeypatch.delenv("RS_PACKED")
    assert (E.BatchedWorld(0, 0, 3, 3, 25, 65536, seed=5).kernel_flags & 8)
    # by world size -- except under RS_OPT_STEP_OVERLAP = 3, the throughput regime, which takes the packed forms at every size
    small = E.BatchedWorld(0, 0, 3, 3, 25, 8192, seed=5)
    assert bool(small.kernel_flags & 8) == (small.get_option(_lib_consts().OPT_STEP_OVERLAP) == 3)
    small.set_option(_lib_consts().OPT_STEP_OVERLAP, 0)
    assert not (small.kernel_flags & 8)
    small.set_option(_lib_consts().OPT_STEP_OVERLAP, 3)
    assert small.kernel_flags & 8
    p.task_reset(E.TASK_VSS_V0); q.task_reset(E.TASK_VSS_V0)
    for _ in range(60):
        op = p.vss_env_step(a, max_steps=25)
        oq = q.vss_env_step(a, max_steps=25)
        assert all(torch.equal(x, y) for x, y in zip(op, oq))
    assert torch.equal(p.get_raw(), q.get_raw())


def test_host_step_equals_the_device_step_at_a_lane_per_match_size(engine):
    """rs_vss_env_step_host (pinned staging, packed D2H) vs the device-tensor step, bit-exact, at
    a world size that runs the lane-per-match kernels with a ragged last warp."""
    from rsoccer_b200 import envs
    n = 20001
    env = envs.make("VSS-v0", num_envs=n, seed=8, max_episode_steps=7)
    ref = envs.make("VSS-v0", num_envs=n, seed=8, max_episode_steps=7)
    env.reset(); ref.reset()
    g = torch.Generator().manual_seed(1)
    for t in range(20):
        a = torch.rand(n, 2, generator=g) * 2 - 1
        o1, r1, d1, t1, _ = env.step(a.cuda())
        o2, r2, d2, t2 = ref.step_host(a.numpy())
        assert np.array_equal(o1.cpu().numpy(), o2) and np.array_equal(r1.cpu().numpy(), r2)
        assert np.array_equal(d1.cpu().numpy(), d2) and np.array_equal(t1.cpu().numpy(), t2)
    assert torch.equal(env.world.get_raw(), ref.world.get_raw()) and env.world.t == ref.world.t


@pytest.mark.parametrize("task", ["vss", "sd", "cp"])
def test_host_step_with_pinned_actions_in_place_or_copied(_engine_module, task):
    """RS_OPT_HOST_COPY_ACTIONS: pinned host actions read in place over PCIe (0), staged with a copy (1) or chosen
    by row size (-1, the default), and pageable actions (always copied) all give the device-tensor step's bits"""
    E, L = _engine_module, _lib_consts()
    kind, ft, nb, ny, tid, ad = {"vss": (0, 0, 3, 3, E.TASK_VSS_V0, 2), "sd": (1, 2, 1, 6, E.TASK_SSL_STATIC_DEFENDERS_V0, 5),
                                 "cp": (1, 2, 1, 1, E.TASK_SSL_CONTESTED_POSSESSION_V0, 5)}[task]
    n = 3001
    g = torch.Generator().manual_seed(5)
    acts = [torch.rand(n, ad, generator=g) * 2 - 1 for _ in range(6)]

    def run(mode, pinned):
        w = E.BatchedWorld(kind, ft, nb, ny, 25, n, seed=4)
        w.task_reset(tid)
        if mode is None:
            outs = []
            for a in acts:
                o = w.vss_env_step(a.cuda()) if task == "vss" else w.ssl_env_step(tid, a.cuda())
                outs.append([x.cpu().clone() for x in o])
            return outs, w.get_raw().cpu()
        assert w.get_option(L.OPT_HOST_COPY_ACTIONS) == -1
        w.set_option(L.OPT_HOST_COPY_ACTIONS, mode)
        h_out = w.alloc_host_outputs(tid)
        h_act = torch.empty(n, ad).pin_memory() if pinned else torch.empty(n, ad)
        outs = []
        for a in acts:
            h_act.copy_(a)
            if task == "vss":
                w.vss_env_step_host(h_act, *h_out)
            else:
                w.ssl_env_step_host(tid, h_act, *h_out)
            outs.append([x.clone() for x in h_out])
        return outs, w.get_raw().cpu()

    ref_outs, ref_raw = run(None, False)
    for mode, pinned in ((-1, True), (0, True), (1, True), (-1, False), (0, False)):
        outs, raw = run(mode, pinned)
        for step, (o, r) in enumerate(zip(outs, ref_outs)):
            for x, y in zip(o, r):
                assert torch.equal(x, y), (mode, pinned, step)
        assert torch.equal(raw, ref_raw), (mode, pinned)
    with pytest.raises(Exception):
        E.BatchedWorld(kind, ft, nb, ny, 25, 8, seed=4).set_option(L.OPT_HOST_COPY_ACTIONS, 2)


def test_split_phase_host_steps_of_two_env_groups(engine):
    """step_async / step_wait (rs_*_env_step_host_begin + rs_host_step_wait): two env groups on two streams,
    one stepping while the other's outputs cross PCIe, give the blocking step_host's bits; a second begin
    while one is pending is refused."""
    from rsoccer_b200 import envs
    n = 4099
    ga = [envs.make(name, num_envs=n, seed=11 + i) for i, name in enumerate(("VSS-v0", "SSLStaticDefenders-v0"))]
    gb = [envs.make(name, num_envs=n, seed=11 + i) for i, name in enumerate(("VSS-v0", "SSLStaticDefenders-v0"))]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for e in ga + gb:
        e.reset()
    torch.cuda.synchronize()
    g = torch.Generator().manual_seed(2)
    acts = [[(torch.rand(n, e.ACT_DIM, generator=g) * 2 - 1).numpy() for e in ga] for _ in range(12)]
    for i, e in enumerate(ga):                      # prologue: both groups in flight
        with torch.cuda.stream(streams[i]):
            e.step_async(acts[0][i])
    with pytest.raises(Exception):
        ga[0].step_async(acts[0][0])
    for t in range(12):
        for i, e in enumerate(ga):
            got = [x.copy() for x in e.step_wait()]
            if t + 1 < 12:
                with torch.cuda.stream(streams[i]):
                    e.step_async(acts[t + 1][i])    # group i steps again while the other group is consumed
            want = gb[i].step_host(acts[t][i])
            for x, y in zip(got, want):
                assert np.array_equal(x, y), (t, i)
    for a, b in zip(ga, gb):
        a.world.host_step_wait()                    # nothing pending: a no-op
        assert torch.equal(a.world.get_raw(), b.world.get_raw())


@pytest.mark.parametrize("env_id", ["VSS-v0", "SSLStaticDefenders-v0", "SSLContestedPossession-v0"])
def test_final_obs_mode_exposes_the_terminal_observation(engine, env_id):
    """final_obs=True: info["final_obs"] is the observation of the step BEFORE any reset (rows flagged in
    info["_final_obs"] = terminated | truncated are terminal observations), obs holds the first observation of the
    next episode for those rows; everything else equals the same env stepped without auto-reset."""
    from rsoccer_b200 import envs
    n, horizon = 700, 5
    a_env = envs.make(env_id, num_envs=n, seed=21, max_episode_steps=horizon, final_obs=True)
    b_env = envs.make(env_id, num_envs=n, seed=21, max_episode_steps=horizon, auto_reset=False)
    o0, _ = a_env.reset()
    o1, _ = b_env.reset()
    assert torch.equal(o0, o1)
    g = torch.Generator().manual_seed(9)
    ended_total = 0
    for t in range(horizon):                      # no env of `b_env` is ever reset, so the two agree until the first end
        act = (torch.rand(n, a_env.ACT_DIM, generator=g) * 2 - 1).cuda()
        oa, ra, da, ta, info = a_env.step(act)
        ob, rb, db, tb, _ = b_env.step(act)
        first = t == 0 or ended_total == 0
        ended = info["_final_obs"]
        assert torch.equal(ended, da | ta)
        if first:
            assert torch.equal(info["final_obs"], ob) and torch.equal(ra, rb) and torch.equal(da, db) and torch.equal(ta, tb)
            keep = ~ended
            assert torch.equal(oa[keep], ob[keep])          # rows that did not end: the ordinary observation
            if ended.any():
                assert not torch.equal(oa[ended], ob[ended])  # ended rows: first observation of the next episode
                assert int(a_env.world.steps[:n][ended].max()) == 0
        ended_total += int(ended.sum())
        if not first:
            break
    assert ended_total > 0                        # the TimeLimit at the latest ends every env within the horizon
    # after the truncation step every row of `a_env` has been re-placed and steps on
    oa, ra, da, ta, info = a_env.step((torch.rand(n, a_env.ACT_DIM, generator=g) * 2 - 1).cuda())
    assert torch.isfinite(oa).all() and int(a_env.world.steps[:n].max()) <= horizon


def test_render_rgb_array_of_env_i_of_a_batch(engine):
    """env.render(index) with render_mode="rgb_array": the picture of ONE match of the batch, drawn from
    the device state (ball where get_state says it is); a window mode is not offered."""
    from rsoccer_b200 import envs, render as RR
    env = envs.make("VSS-v0", num_envs=5, seed=2, render_mode="rgb_array")
    env.reset()
    img = env.render(3)
    st = env.world.get_state()[3].cpu().numpy()
    f = env.world.field_params()
    m = 0.1 + f["goal_depth"]
    sc = 750 / (f["length"] + 2 * m)
    assert img.shape[1] == 750 and img.dtype == np.uint8
    assert tuple(img[int((f["width"] / 2 + m - st[1]) * sc), int((st[0] + f["length"] / 2 + m) * sc)]) == RR.ORANGE
    with pytest.raises(NotImplementedError):
        envs.make("VSS-v0", num_envs=1, render_mode="human")
    ssl = envs.make("SSLStaticDefenders-v0", num_envs=2, render_mode="rgb_array")
    ssl.reset()
    assert ssl.render(1, width_px=400).shape[1] == 400


def _twin_worlds(E, n, seed, modes):
    ws = []
    for m in modes:
        w = E.BatchedWorld(0, 0, 3, 3, 25, n, seed=seed)
        w.set_option(_lib_consts().OPT_STEP_OVERLAP, m)
        w.task_reset(E.TASK_VSS_V0)
        ws.append(w)
    return ws


def _lib_consts():
    from rsoccer_b200 import _lib
    return _lib


@pytest.mark.parametrize("n", [65536, 5000, 33])
def test_step_overlap_modes_are_bit_identical(_engine_module, monkeypatch, n):
    """RS_OPT_STEP_OVERLAP: consecutive steps synchronise per 32-match tile instead of grid-wide.
    Same kernels, same arithmetic, only the waiting differs -- so after hundreds of back-to-back
    steps (direct launches and CUDA-graph replays, nothing enqueued in between, fixed action
    buffer) every mode must leave the world in the same state bit for bit; a tile that started on
    stale state would diverge chaotically.  33 and 5 000 matches leave ragged tiles / dead warps."""
    E, L = _engine_module, _lib_consts()
    monkeypatch.setenv("RS_PER_MATCH", "1")
    monkeypatch.delenv("RS_STEP_OVERLAP", raising=False)
    ws = _twin_worlds(E, n, 11, (0, 1, 2, 3))
    assert [w.get_option(L.OPT_STEP_OVERLAP) for w in ws] == [0, 1, 2, 3]
    a = (torch.rand(n, 2, device="cuda") * 2 - 1)
    outs = [w.alloc_outputs(E.TASK_VSS_V0) for w in ws]
    s = torch.cuda.Stream()
    finals = []
    for w, out in zip(ws, outs):
        with torch.cuda.stream(s):
            for _ in range(150):
                w.vss_env_step(a, out=out, max_steps=40)
            s.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                for _ in range(25):
                    w.vss_env_step(a, out=out, max_steps=40)
            for _ in range(8):
                g.replay()
            s.synchronize()
        assert w.get_option(L.OPT_OVERLAP_ERRORS) == 0
        finals.append((w.state.clone(), [o.clone() for o in out], w.sync_t()))
    assert finals[0][2] == 150 + 200
    for f in finals[1:]:
        assert f[2] == finals[0][2]
        assert torch.equal(f[0], finals[0][0])
        assert all(torch.equal(x, y) for x, y in zip(f[1], finals[0][1]))


@pytest.mark.parametrize("task,per_match,n", [("vss", 0, 4096), ("sd", 0, 4096), ("sd", 1, 20000), ("cp", 0, 5000), ("cp", 1, 16384),
                                              ("drib", 1, 3000), ("pass", 1, 3000)])
def test_step_overlap_in_every_task_kernel(_engine_module, monkeypatch, task, per_match, n):
    """the tile protocol in the other step kernels (one lane per body: a tile is the 32 / L matches of a warp; the
    SSL tasks): chained launches from a replayed graph against the serialised twin, bit for bit"""
    E, L = _engine_module, _lib_consts()
    monkeypatch.setenv("RS_PER_MATCH", str(per_match))
    monkeypatch.delenv("RS_STEP_OVERLAP", raising=False)
    spec = {"vss": (0, 0, 3, 3, E.TASK_VSS_V0, 2), "sd": (1, 2, 1, 6, E.TASK_SSL_STATIC_DEFENDERS_V0, 5),
            "cp": (1, 2, 1, 1, E.TASK_SSL_CONTESTED_POSSESSION_V0, 5), "drib": (1, 2, 1, 4, E.TASK_SSL_DRIBBLING_V0, 4),
            "pass": (1, 2, 2, 0, E.TASK_SSL_PASS_ENDURANCE_V0, 3)}[task]
    kind, ft, nb, ny, tid, ad = spec
    a = torch.rand(n, ad, device="cuda") * 2 - 1
    s = torch.cuda.Stream()
    finals = []
    for mode in (0, 2):
        w = E.BatchedWorld(kind, ft, nb, ny, 25, n, seed=31)
        w.set_option(L.OPT_STEP_OVERLAP, mode)
        w.task_reset(tid)
        out = w.alloc_outputs(tid)

        def step():
            if task == "vss":
                w.vss_env_step(a, out=out, max_steps=60)
            else:
                w.ssl_env_step(tid, a, out=out, max_steps=60)
        with torch.cuda.stream(s):
            for _ in range(20):
                step()
            s.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                for _ in range(20):
                    step()
            for _ in range(10):
                g.replay()
            s.synchronize()
        assert w.get_option(L.OPT_OVERLAP_ERRORS) == 0
        finals.append((w.state.clone(), [o.clone() for o in out]))
    assert torch.equal(finals[0][0], finals[1][0])
    assert all(torch.equal(x, y) for x, y in zip(finals[0][1], finals[1][1]))


def test_step_overlap_with_worlds_in_rotation(_engine_module, monkeypatch):
    """the benchmark's pattern: several worlds stepped round-robin on one stream from one captured graph, so that
    consecutive launches are independent and overlap freely (mode 3: the dense build); every world must end
    exactly where its serialised twin ends"""
    E, L = _engine_module, _lib_consts()
    monkeypatch.setenv("RS_PER_MATCH", "1")
    n, M = 40000, 3
    a = [(torch.rand(n, 2, device="cuda") * 2 - 1) for _ in range(M)]
    s = torch.cuda.Stream()
    finals = []
    for mode in (0, 3):
        ws = []
        for m in range(M):
            w = E.BatchedWorld(0, 0, 3, 3, 25, n, seed=21, env_offset=m * n)
            w.set_option(L.OPT_STEP_OVERLAP, mode)
            w.task_reset(E.TASK_VSS_V0)
            ws.append(w)
        outs = [w.alloc_outputs(E.TASK_VSS_V0) for w in ws]
        with torch.cuda.stream(s):
            for i in range(3 * M):
                ws[i % M].vss_env_step(a[i % M], out=outs[i % M], max_steps=50)
            s.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                for i in range(4 * M):
                    ws[i % M].vss_env_step(a[i % M], out=outs[i % M], max_steps=50)
            for _ in range(20):
                g.replay()
            s.synchronize()
        assert all(w.get_option(L.OPT_OVERLAP_ERRORS) == 0 for w in ws)
        finals.append([w.state.clone() for w in ws])
    assert all(torch.equal(x, y) for x, y in zip(*finals))


def test_step_overlap_survives_interleaved_calls(_engine_module, monkeypatch):
    """anything else that touches the state between two steps (reset, set_raw, rs_step, a masked task
    reset, an option change) puts the next step back on the grid-wide wait: same results as mode 0"""
    E, L = _engine_module, _lib_consts()
    monkeypatch.setenv("RS_PER_MATCH", "1")
    n = 4096
    w0, w2 = _twin_worlds(E, n, 5, (0, 2))
    a = (torch.rand(n, 2, device="cuda") * 2 - 1)
    cmds = torch.rand(n, 6, 2, device="cuda") * 20 - 10
    mask = (torch.arange(n, device="cuda") % 3 == 0)
    for w in (w0, w2):
        for i in range(60):
            w.vss_env_step(a)
            if i % 7 == 3:
                w.step(cmds)
            if i % 11 == 5:
                w.task_reset(E.TASK_VSS_V0, mask=mask)
            if i % 13 == 6:
                w.set_raw(w.get_raw())
            if i % 17 == 8:
                w.state_written()
        torch.cuda.synchronize()
    assert w2.get_option(L.OPT_OVERLAP_ERRORS) == 0
    assert torch.equal(w0.state, w2.state) and w0.sync_t() == w2.sync_t()


def test_rs_step_between_graph_replays_does_not_rewind_the_noise(_engine_module, monkeypatch):
    """rs_step draws no random numbers and must leave the device-authoritative step counter alone:
    graph replays, then rs_step, then more replays == the same sequence launched eagerly"""
    E = _engine_module
    monkeypatch.setenv("RS_PER_MATCH", "1")
    n = 512
    a = torch.rand(n, 2, device="cuda") * 2 - 1
    cmds = torch.rand(n, 6, 2, device="cuda") * 20 - 10
    w = E.BatchedWorld(0, 0, 3, 3, 25, n, seed=4)
    w.task_reset(E.TASK_VSS_V0)
    out = w.alloc_outputs(E.TASK_VSS_V0)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        w.vss_env_step(a, out=out)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            w.vss_env_step(a, out=out)
        for _ in range(3):
            g.replay()
        w.step(cmds)
        for _ in range(2):
            g.replay()
        s.synchronize()
    assert w.sync_t() == 6
    ref = E.BatchedWorld(0, 0, 3, 3, 25, n, seed=4)
    ref.task_reset(E.TASK_VSS_V0)
    for _ in range(4):
        ref.vss_env_step(a)
    ref.step(cmds)
    for _ in range(2):
        ref.vss_env_step(a)
    assert ref.t == 6 and torch.equal(ref.state, w.state)
    # a pending rs_set_t must not be baked into a graph
    w.t = 3
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with pytest.raises(Exception):
            with torch.cuda.graph(g2, stream=s):
                w.vss_env_step(a, out=out)
    with pytest.raises(Exception):
        w.t = 1 << 32


def test_options_are_per_handle(_engine_module, monkeypatch):
    """no process-global switches: two handles created under different RS_PDL settings coexist and agree"""
    E, L = _engine_module, _lib_consts()
    n = 2048
    monkeypatch.setenv("RS_PDL", "0")
    w0 = E.BatchedWorld(0, 0, 3, 3, 25, n, seed=2)
    monkeypatch.setenv("RS_PDL", "1")
    w1 = E.BatchedWorld(0, 0, 3, 3, 25, n, seed=2)
    monkeypatch.delenv("RS_PDL")
    assert (w0.get_option(L.OPT_PDL), w1.get_option(L.OPT_PDL)) == (0, 1)
    a = torch.rand(n, 2, device="cuda") * 2 - 1
    for w in (w0, w1):
        w.task_reset(E.TASK_VSS_V0)
    for _ in range(20):
        w0.vss_env_step(a); w1.vss_env_step(a)
    assert torch.equal(w0.state, w1.state)
    w1.set_option(L.OPT_PDL, 0)
    assert w1.get_option(L.OPT_PDL) == 0 and w0.get_option(L.OPT_PDL) == 0
    with pytest.raises(Exception):
        w0.set_option(99, 1)


def test_world_runs_on_its_own_device_whatever_is_current(_engine_module):
    """every entry point selects the world's device and restores the caller's (two worlds on two GPUs
    in one process; device="cuda" resolves to the current device)"""
    E = _engine_module
    assert E.BatchedWorld(0, 0, 3, 3, 25, 8, device="cuda").device.index == torch.cuda.current_device()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    n = 1024
    a = torch.rand(n, 2) * 2 - 1
    w0 = E.BatchedWorld(0, 0, 3, 3, 25, n, device="cuda:0", seed=6)
    w1 = E.BatchedWorld(0, 0, 3, 3, 25, n, device="cuda:1", seed=6)
    a0, a1 = a.to("cuda:0"), a.to("cuda:1")
    with torch.cuda.device(0):                  # cuda:0 stays current while cuda:1's world is stepped
        for w in (w0, w1):
            w.task_reset(E.TASK_VSS_V0)
        for _ in range(10):
            w0.vss_env_step(a0); w1.vss_env_step(a1)
        h0, h1 = w0.alloc_host_outputs(E.TASK_VSS_V0), w1.alloc_host_outputs(E.TASK_VSS_V0)
        ha = a.pin_memory()
        w0.vss_env_step_host(ha, *h0); w1.vss_env_step_host(ha, *h1)
        assert torch.cuda.current_device() == 0
    assert torch.equal(w0.state.cpu(), w1.state.cpu()) and torch.equal(h0[0], h1[0])

"""Host-side logic on CPU: Frame parsing, the gymnasium stand-in, env sharding + the rollout
all-gather on a world_size-2 gloo group."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def _vss_state(n=None):
    rng = np.random.default_rng(0)
    return rng.normal(size=(5 + 6 * 6,) if n is None else (n, 5 + 6 * 6))


def test_frame_parse_flat_is_scalar_like_the_reference():
    from rsoccer_b200.entities import FrameSSL, FrameVSS
    st = _vss_state()
    f = FrameVSS().parse(list(st), 3, 3)
    assert f.ball.x == st[0] and f.ball.v_y == st[4]
    assert f.robots_blue[2].theta == st[5 + 12 + 2] and f.robots_yellow[0].x == st[5 + 18]
    assert sorted(f.robots_blue) == [0, 1, 2] and f.robots_yellow[1].id == 1
    st = np.random.default_rng(1).normal(size=5 + 11 * 3)
    st[5 + 6] = 1.0
    g = FrameSSL().parse(st, 1, 2)
    assert g.robots_blue[0].infrared is True and g.robots_yellow[1].v_wheel3 == st[5 + 22 + 10]


def test_frame_parse_batched_gives_column_views():
    from rsoccer_b200.entities import FrameVSS
    st = torch.tensor(_vss_state(7))
    f = FrameVSS().parse(st, 3, 3)
    assert f.ball.x.shape == (7,) and torch.equal(f.robots_yellow[2].v_theta, st[:, 5 + 30 + 5])
    st[:, 0] = 42.0
    assert (f.ball.x == 42.0).all()          # a view, not a copy


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_frame_parse_equals_reference_parser():
    """same flat state through rsoccer_gym/Entities/Frame.py and through ours"""
    import importlib.util
    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m); return m
    sys.path.insert(0, REF)
    try:
        ents = {}
        for n in ("Ball", "Robot"):
            ents[n] = load("rsoccer_gym.Entities." + n, os.path.join(REF, "rsoccer_gym", "Entities", n + ".py"))
            sys.modules["rsoccer_gym.Entities." + n] = ents[n]
        ref = load("ref_frame", os.path.join(REF, "rsoccer_gym", "Entities", "Frame.py"))
    finally:
        sys.path.remove(REF)
        for n in ("Ball", "Robot"):
            sys.modules.pop("rsoccer_gym.Entities." + n, None)
    from rsoccer_b200.entities import FrameSSL, FrameVSS
    st = _vss_state()
    a, b = ref.FrameVSS(), FrameVSS()
    a.parse(st, 3, 3); b.parse(st, 3, 3)
    st2 = np.random.default_rng(3).normal(size=5 + 11 * 7)
    c, d = ref.FrameSSL(), FrameSSL()
    c.parse(st2, 1, 6); d.parse(st2, 1, 6)
    for x, y in ((a, b), (c, d)):
        assert vars(x.ball) == vars(y.ball)
        for team in ("robots_blue", "robots_yellow"):
            assert getattr(x, team).keys() == getattr(y, team).keys()
            for k in getattr(x, team):
                assert vars(getattr(x, team)[k]) == vars(getattr(y, team)[k])


def test_gymnasium_stand_in_make_register_timelimit():
    from rsoccer_b200.compat import gymnasium_shim as gym

    class Dummy(gym.Env):
        def __init__(self, k=1):
            self.k = k
            self.action_space = gym.spaces.Box(low=-1, high=1, shape=(2,), dtype=np.float32)

        def reset(self, *, seed=None, options=None):
            super().reset(seed=seed)
            return np.zeros(1), {}

        def step(self, a):
            return np.zeros(1), 0.0, False, False, {}

    gym.register(id="Dummy-v0", entry_point=Dummy, max_episode_steps=3, kwargs={"k": 5})
    env = gym.make("Dummy-v0")
    assert env.unwrapped.k == 5 and env.action_space.low.shape == (2,) and env.action_space.high[0] == 1
    env.reset(seed=1)
    tr = [env.step(env.action_space.sample())[3] for _ in range(3)]
    assert tr == [False, False, True]
    env.reset()
    assert env.step(None)[3] is False


def test_shard_range_partitions_everything():
    from rsoccer_b200.sharding import shard_range
    for n, w in ((262144, 8), (65536, 3), (5, 8), (0, 2)):
        r = [shard_range(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) <= 1
    assert [shard_range(262144, k, 8)[1] - shard_range(262144, k, 8)[0] for k in range(8)] == [32768] * 8
    with pytest.raises(ValueError):
        shard_range(10, 4, 4)


def _gloo_worker(rank, world, port, n_total, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rsoccer_b200.sharding import gather_rollout, shard_range
    lo, hi = shard_range(n_total, rank, world)
    # a rollout tensor whose content encodes the GLOBAL env id
    obs = torch.arange(lo, hi, dtype=torch.float32)[:, None].repeat(1, 4)
    full = gather_rollout(obs, n_total=n_total)
    traj = torch.arange(lo, hi, dtype=torch.float32)[None, :, None].repeat(3, 1, 2)     # [T, n_local, 2]
    full_t = gather_rollout(traj, n_total=n_total, dim=1)
    ok = bool(torch.equal(full[:, 0], torch.arange(n_total, dtype=torch.float32)) and full.shape == (n_total, 4)
              and full_t.shape == (3, n_total, 2)
              and torch.equal(full_t[1, :, 1], torch.arange(n_total, dtype=torch.float32)))
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 7])
def test_rollout_all_gather_world_size_2_gloo(n_total):
    """the only collective of the design, on CPU tensors over gloo"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + n_total
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def _pixel(img, field, x, y, width_px=750, margin_m=0.1):
    """colour at field point (x, y) [m]; same mapping as rsoccer_b200/render.py"""
    m = margin_m + field["goal_depth"]
    sc = width_px / (field["length"] + 2 * m)
    return tuple(int(v) for v in img[int((field["width"] / 2 + m - y) * sc), int((x + field["length"] / 2 + m) * sc)])


@pytest.mark.parametrize("kind", ["vss", "ssl"])
def test_render_rgb_array_of_one_match(oracle, kind):
    """rgb_array picture of one match (stands in for Render/ + pygame, vss_gym_base.py:148-187): shape,
    palette (Render/utils.py:2-15), ball / robots / lines where the state says they are."""
    from rsoccer_b200 import render as RR
    O = oracle
    if kind == "vss":
        w = O.OracleWorld(O.KIND_VSS, 0, 3, 3, 25, 1)
        nb, ny, K = 3, 3, 6
    else:
        w = O.OracleWorld(O.KIND_SSL, 2, 1, 2, 25, 1)
        nb, ny, K = 1, 2, 11
    field = w.field_params()
    L, W = field["length"], field["width"]
    st = np.zeros(5 + K * (nb + ny))
    st[0:2] = (0.25 * L, 0.2 * W)                                    # ball
    pos = [(-0.3 * L, 0.25 * W, 0.0), (-0.2 * L, -0.3 * W, 90.0), (0.1 * L, -0.1 * W, 200.0),
           (0.3 * L, -0.3 * W, 180.0), (0.35 * L, 0.3 * W, -45.0), (0.0, 0.35 * W, 10.0)][:nb + ny]
    for k, (x, y, th) in enumerate(pos):
        st[5 + K * k: 8 + K * k] = (x, y, th)
    img = RR.render_rgb(st, field, kind, nb, ny)
    assert img.dtype == np.uint8 and img.shape[1] == 750 and img.shape[2] == 3
    assert abs(img.shape[0] / img.shape[1] - (W + 2 * (0.1 + field["goal_depth"])) / (L + 2 * (0.1 + field["goal_depth"]))) < 0.01
    assert _pixel(img, field, st[0], st[1]) == RR.ORANGE
    assert _pixel(img, field, 0.45 * L, 0.45 * W) == RR.BG_GREEN
    assert _pixel(img, field, 0.0, -0.4 * W) == RR.WHITE             # centre line
    assert _pixel(img, field, L / 2, 0.4 * W) == RR.WHITE            # field outline
    r = field["rbt_radius"]
    for k, (x, y, th) in enumerate(pos):
        c, s = np.cos(np.radians(th)), np.sin(np.radians(th))
        # a point ahead-left of the centre lies inside the team-coloured tag, off the heading mark
        px = _pixel(img, field, x + 0.3 * r * c - 0.3 * r * s, y + 0.3 * r * s + 0.3 * r * c)
        assert px == (RR.BLUE if k < nb else RR.YELLOW), (k, px)
    with pytest.raises(ValueError):
        RR.render_rgb(st[:-1], field, kind, nb, ny)


def test_committed_bench_lines_keep_the_driver_contract():
    """the JSON lines bench.py printed on the B200 (profiles/r2_bench_line*.json) carry every key the driver and the
    tier's measurement section name, with consistent values"""
    import json
    prof = os.path.join(ROOT, "profiles")
    line = json.load(open(os.path.join(prof, "r2_bench_line.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in line, k
    assert line["n_gpus"] == 1 and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert "workload" in line["config"] and "model" not in line["config"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(line["clocks"])
    e = line["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e)
    assert e["h2d_bytes_per_step"] == 65536 * 2 * 4 and e["d2h_bytes_per_step"] == 65536 * (40 * 4 + 4 + 1 + 1)
    assert 0 < e["value"] < line["value"]                     # host copies inside the timed region: never the device figure
    r = line["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    # achieved = 592 algorithmic bytes x 65 536 matches / measured launch time
    assert abs(r["achieved"] - 592 * 65536 / (line["ms_per_step"] * 1e-3) / 1e9) < 1e-3 * r["achieved"]
    assert abs(line["value"] - 65536 / (line["ms_per_step"] * 1e-3)) < 1e-6 * line["value"]
    c = line["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] in ("port", "reference")
    assert line["gpu_launches"] > 0
    for name in ("vss4096", "sd4096", "cp16384", "vss262144_sharded"):
        assert {"ms_per_step", "roofline", "e2e"} <= set(line["configs"][name]), name
    assert line["strong"]["scaling"] == "strong" and line["strong"]["envs_total"] == 65536
    ref = json.load(open(os.path.join(prof, "r2_bench_line_reference.json")))
    assert ref["impl"] == "reference" and ref["metric"] == line["metric"] and ref["unit"] == line["unit"]
    assert ref["cpu_baseline"]["value"] == ref["value"] == ref["e2e"]["value"]
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["e2e"]["d2h_bytes_per_step"] == 0
    for n in (2, 4, 8):
        ln = json.load(open(os.path.join(prof, "r2_bench_line_n%d.json" % n)))
        assert ln["n_gpus"] == n and ln["scaling"] == "weak" and ln["config"]["envs_per_gpu"] == 65536
        assert abs(ln["value"] - n * 65536 / (ln["ms_per_step"] * 1e-3)) < 1e-6 * ln["value"]

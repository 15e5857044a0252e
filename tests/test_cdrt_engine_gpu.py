"""GPU: the CDRT-RRT* engine (rtpp_cdrt_*, one warp per BImap planner) driven through the scripted demo loop of
CDRT2/test/testD.cpp on virtual time, against (a) the golden vectors produced by the reference's own CDRT.cpp
(tests/golden/cdrt_*.npz) and (b) the CPU oracle live on another map.  Node sets, parents, costs (binary64 bits),
invM, children counts, paths, path costs, queue length, rand/clock positions: all bit-identical, tick after tick."""
import os

import numpy as np
import pytest

import cdrt_scenarios as sc

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def gpu_row(pl, path, col):
    st = pl.state()
    d = pl.dump()
    return d, (len(d[0]), st["queue"], st["rand_pos"], st["clock_reads"], st["island"], st["n_trees"], st["samples_left"], int(col), len(path))


@pytest.mark.parametrize("case", sc.GOLDEN_CASES, ids=lambda c: c["name"])
def test_engine_reproduces_reference_golden(ctx, oracle, case):
    import rtpp_b200

    z = np.load(os.path.join(HERE, "golden", f"cdrt_{case['name']}.npz"))
    fx = oracle.load_cdrt_fixture(case["map"])
    pl = rtpp_b200.CdrtPlanner(ctx, fx, queue_capacity=1 << 21)
    sc.setup(pl, case["map"], case["goal"], oracle.rand31(case["seed"], case["n_rand"]), dt_ns=case["dt"], params=case["params"])
    for i, a in enumerate(pl.dump()):
        assert np.array_equal(a.view(np.uint8), z[f"start_dump{i}"].view(np.uint8)), f"start dump {i}"
    for k in range(case["ticks"]):
        path, col, cost = sc.tick(pl, case["map"], k)
        assert pl.state()["status"] == 0, pl.state()
        d, got = gpu_row(pl, path, col)
        r = z["rows"][k]
        want = tuple(int(v) for v in (r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[9], r[10]))
        assert got == want, f"tick {k}: {got} vs {want}"
        assert np.array_equal(path, z[f"t{k}_path"]), f"tick {k} path"
        assert np.float64(cost).tobytes() == z[f"t{k}_cost"].tobytes(), f"tick {k} path cost"
        if k in case["dump_ticks"]:
            for i, a in enumerate(d):
                assert np.array_equal(a.view(np.uint8), z[f"t{k}_dump{i}"].view(np.uint8)), f"tick {k} dump {i}"
            has, cnt = pl.eng.read_islands()
            assert np.array_equal(np.flatnonzero(has), z[f"t{k}_islands"]) and np.array_equal(cnt[has != 0], z[f"t{k}_island_pts"])
        assert np.array_equal(sc.dump_hash(d), z[f"t{k}_hash"]), f"tick {k} node dump hash"
    pl.close()


@pytest.mark.parametrize("name,goal,seed,dt,ticks", [("GAME0", 2, 3, 5000, 25), ("MAZE0", 1, 5, 20000, 20)])
def test_engine_vs_oracle_live(ctx, oracle, name, goal, seed, dt, ticks):
    import rtpp_b200

    fx = oracle.load_cdrt_fixture(name)
    r = oracle.rand31(seed, 600_000)
    o = oracle.CdrtCpu(fx, "oracle")
    pl = rtpp_b200.CdrtPlanner(ctx, fx)
    for p in (o, pl):
        sc.setup(p, name, goal, r, dt_ns=dt)
    for k in range(ticks):
        pa, pb = sc.tick(o, name, k), sc.tick(pl, name, k)
        assert pl.state()["status"] == 0, pl.state()
        assert np.array_equal(pa[0], pb[0]) and pa[1] == pb[1] and np.float64(pa[2]).tobytes() == np.float64(pb[2]).tobytes(), f"tick {k} path"
        da, db = o.dump(), pl.dump()
        assert len(da[0]) == len(db[0]), f"tick {k}"
        for x, y in zip(da, db):
            assert np.array_equal(x.view(np.uint8), y.view(np.uint8)), f"tick {k}"
        so, sg = o.state(), pl.state()
        assert (so["queue"], o.rand_pos(), o.clock_reads(), so["island"], so["n_trees"], so["samples_left"]) == \
               (sg["queue"], sg["rand_pos"], sg["clock_reads"], sg["island"], sg["n_trees"], sg["samples_left"]), f"tick {k}"
    c = o.counters()
    assert (c["pops"], c["rewired"], c["sampled"], c["pushes"]) == (sg["pops"], sg["rewired"], sg["sampled"], sg["pushes"])
    o.close()
    pl.close()


def _engine(ctx, fx, n=1, **kw):
    import rtpp_b200

    return rtpp_b200.CdrtEngine(ctx, n, fx["graphs"][0], fx["amap"], fx["imap"], fx["resolution"], **kw)


def test_two_instances_are_independent(ctx, oracle):
    """Two planners of one engine (different goals, seeds and obstacle sets) each follow their own oracle run."""
    import rtpp_b200

    name = "MESS6"
    fx = oracle.load_cdrt_fixture(name)
    eng = _engine(ctx, fx, n=2, rand_capacity=200_000)
    start, goals = sc.TASKS[name]
    orcs = []
    for i in range(2):
        o = oracle.CdrtCpu(fx, "oracle")
        o.set_params()
        o.set_clock(10000)
        r = oracle.rand31(40 + i, 200_000)
        o.set_rand(r)
        obs = sc.OBSTACLES[name][: 4 - 2 * i]  # instance 1 has only two obstacles
        for k, (typ, x, y, ang, rad, outline) in enumerate(obs):
            o.add_obstacle(k, typ, float(x), float(y), float(ang), float(rad), np.array(outline, np.float64), 2.0)
            eng.add_obstacle(i, k, typ, float(x), float(y), float(ang), float(rad), np.array(outline, np.float64), 2.0)
        for p in (o, None):
            if p is not None:
                p.set_agent(float(start[0]), float(start[1]))
                p.set_goal(*[float(v) for v in goals[i + 1]])
        eng.set_agent(i, float(start[0]), float(start[1]))
        eng.set_goal(i, *[float(v) for v in goals[i + 1]])
        o.start()
        eng.start(i, r)
        orcs.append((o, len(obs)))
    for k in range(12):
        eng.tick(10000, get_path=True)
        t = (k + 1) * sc.TICK_S
        for i, (o, nobs) in enumerate(orcs):
            o.task()
            po = o.get_path()
            pl = rtpp_b200.CdrtPlanner(ctx, fx, inst=i, engine=eng)
            pg = pl.get_path()
            assert np.array_equal(po[0], pg[0]) and po[1] == pg[1] and np.float64(po[2]).tobytes() == np.float64(pg[2]).tobytes()
            for a, b in zip(o.dump(), pl.dump()):
                assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), f"tick {k} instance {i}"
            for j in range(nobs):
                q = sc.along(sc.OBS_PATHS[name][j], sc.OBS_SPEED * t)
                o.move_obstacle(j, float(q[0]), float(q[1]), 0.0)
                eng.move_obstacle(i, j, float(q[0]), float(q[1]), 0.0)
    assert (eng.stats()[:, 11] == 0).all()
    for o, _ in orcs:
        o.close()
    eng.close()


def test_no_goal_no_obstacles_and_rejected_points(ctx, oracle):
    """GetPath without a goal returns {true, doubleMax} and an empty path (CDRT.cpp:1069-1075); with no obstacle no island
    ever holds a kd-tree, so every sample is a cut-line sample; points outside the free polygons are ignored
    (RTRT_SetAgent/SetGoal print and return, CDRT.cpp:384-423)."""
    fx = oracle.load_cdrt_fixture("MESS6")
    eng = _engine(ctx, fx, rand_capacity=100_000)
    o = oracle.CdrtCpu(fx, "oracle")
    r = oracle.rand31(3, 100_000)
    o.set_params()
    o.set_clock(20000)
    o.set_rand(r)
    assert eng.set_agent(0, 5.0, 5.0) == 1          # inside an obstacle: rejected
    assert eng.set_agent(0, 1e6, 10.0) == 1         # outside the map: rejected
    assert eng.set_goal(0, 5.0, 5.0) == 1
    assert eng.set_agent(0, 720.0, 720.0) == 0
    o.set_agent(720.0, 720.0)
    o.start()
    eng.start(0, r)
    for _ in range(5):
        o.task()
        eng.tick(20000, get_path=True)
    s = eng.stats_dict(0)
    assert s["status"] == 0 and s["n_trees"] == 0 and s["path_len"] == 0 and s["path_col"] == 1
    assert np.array(s["path_cost_bits"], np.int64).view(np.float64) == 1.79769e307
    so = o.state()
    assert (s["queue"], s["samples_left"], s["rand_pos"], s["clock_reads"]) == (so["queue"], so["samples_left"], o.rand_pos(), o.clock_reads())
    o.close()
    eng.close()


def test_capacity_overflows_are_reported_not_fatal(ctx, oracle):
    """A full queue or node store stops the instance with a status code (2 / 1); nothing is written out of bounds."""
    fx = oracle.load_cdrt_fixture("MESS6")
    r = oracle.rand31(7, 400_000)
    for kw, want in ((dict(queue_capacity=512), 2), (dict(node_capacity=16), 0)):
        eng = _engine(ctx, fx, rand_capacity=400_000, **kw)
        for k, (typ, x, y, ang, rad, outline) in enumerate(sc.OBSTACLES["MESS6"]):
            eng.add_obstacle(0, k, typ, float(x), float(y), float(ang), float(rad), np.array(outline, np.float64), 2.0)
        eng.set_agent(0, 720.0, 720.0)
        eng.set_goal(0, 70.0, 70.0)
        eng.start(0, r)
        for _ in range(14):
            eng.tick(20000)
        st = eng.stats_dict(0)
        if want:
            assert st["status"] == want, st
        else:  # the node store is sized up to the cut-line samples + 64: island samples then overflow it
            assert st["status"] in (0, 1), st
        eng.close()

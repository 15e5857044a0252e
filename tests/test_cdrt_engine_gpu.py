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

"""RT-FMT engine parity (F1/F2/R3/P1): the GPU engine against (a) the oracle run side by side and
(b) the golden vectors produced by the REFERENCE's own RTFMT.cpp (tests/golden/rtfmt_*.npz): parents,
cost bits, collision flags, set membership (V_unv/V_open/V_opennew/V_closed/V_toopen), z and the path."""
import os

import numpy as np
import pytest

from helpers import dyn_plane, golden_cases

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("case", golden_cases()["rtfmt"], ids=lambda c: c["name"])
def test_fmt_engine_reproduces_reference_golden(ctx, oracle, case):
    import rtpp_b200

    g = np.load(os.path.join(GOLD, f"rtfmt_{case['name']}.npz"))
    S = oracle.load_map(case["map"])
    H, W = S.shape
    orc = oracle.OrcRTFMT(S)  # the host-side sampler of Start(): gives the node set
    assert orc.set_agent(*case["agent"]) and orc.set_goal(*case["goal"])
    orc.start(oracle.rand31(case["seed"], case["n_rand"]), case["samples"])
    xy = orc.dump()[0]
    assert np.array_equal(xy, g["xy"])
    eng = rtpp_b200.FmtEngine(ctx, 1, 4096, 5, W, H)
    eng.set_map(0, 1, S)
    eng.start(np.zeros(1, np.int32), [xy], np.array([case["goal"]], np.float32))
    rng = np.random.default_rng(case["seed"])
    fallbacks = 0
    for rnd in range(case["rounds"]):
        if rnd in case["dobs_rounds"]:
            D = dyn_plane(S.shape, rng, n_blobs=5)
            eng.set_map(0, 2, D)
            orc.set_dobs(D)
        eng.tick(update_tree=True, iters=500, with_rewire=True, get_path=True)
        orc.update_tree()
        orc.expand(500, True)
        par, c, cl, st = eng.read_tree(0)
        d = eng.stats_dict(0)
        k = f"r{rnd}_"
        assert d["status"] == 0
        # vs the reference's golden vectors
        assert np.array_equal(par, g[k + "par"]), f"round {rnd}"
        assert np.array_equal(c.view(np.uint32), g[k + "cost"].view(np.uint32))
        assert np.array_equal(cl, g[k + "col"]) and np.array_equal(st, g[k + "state"])
        zref = g[k + "z"]
        assert (d["z"] < 0 and zref[0] < 0) or np.array_equal(xy[d["z"]], zref)
        # vs the oracle run side by side (also covers the counters)
        oxy, opar, oc, ocl, ost = orc.dump()
        assert np.array_equal(par, opar) and np.array_equal(st, ost)
        # GetPath, both branches on the device: path_ok = 1 ends in the goal, 0 is the depth-5 fallback chain (RTFMT.cpp:322-391)
        gpath = np.asarray(g[k + "path"], np.float32).reshape(-1, 2)
        ids = eng.read_path(0)
        p = np.array([case["goal"] if j == -2 else xy[j] for j in ids], np.float32).reshape(-1, 2)
        assert np.array_equal(p, gpath), f"round {rnd}: path (path_ok {d['path_ok']})"
        fallbacks += int(not d["path_ok"] and len(ids) > 0)
        if rnd in case["move_rounds"]:
            p = orc.get_path()
            if len(p) > 2:
                orc.set_agent(float(p[1][0]), float(p[1][1]))
                eng.set_inputs(agent_xy=np.array([p[1]], np.float32))
    assert d["n_unv"] < (400 if case["name"] == "mess6" else 1000)  # the expansion really ran
    eng.close()


def test_fmt_engine_many_instances_against_oracle(ctx, oracle):
    import rtpp_b200

    S = oracle.load_map("Trap")
    H, W = S.shape
    n_inst = 6
    eng = rtpp_b200.FmtEngine(ctx, n_inst, 2048, 5, W, H)
    eng.set_map(0, 1, S)
    orcs, nodes = [], []
    for i in range(n_inst):
        o = oracle.OrcRTFMT(S)
        assert o.set_agent(660.0, 660.0) and o.set_goal(1160.0, 130.0)
        o.start(oracle.rand31(50 + i, 50000), 1000 + 100 * i)
        orcs.append(o)
        nodes.append(o.dump()[0])
    eng.start(np.zeros(n_inst, np.int32), nodes, np.tile(np.array([[1160.0, 130.0]], np.float32), (n_inst, 1)))
    rng = np.random.default_rng(1)
    for rnd in range(5):
        if rnd == 2:
            D = dyn_plane(S.shape, rng, n_blobs=5)
            eng.set_map(0, 2, D)
            for o in orcs:
                o.set_dobs(D)
        eng.tick(update_tree=True, iters=400, with_rewire=True, get_path=False)
        for i, o in enumerate(orcs):
            o.update_tree()
            o.expand(400, True)
            oxy, opar, oc, ocl, ost = o.dump()
            par, c, cl, st = eng.read_tree(i)
            assert np.array_equal(par, opar) and np.array_equal(st, ost) and np.array_equal(cl, ocl)
            assert np.array_equal(c.view(np.uint32), oc.view(np.uint32))
    eng.close()

"""Drop-in parity: the C++ RTRRTplanner class (rtpp_b200/planners, GPU engine underneath) driven
through the reference's own method sequence must reproduce the golden vectors that the REFERENCE
produced (tests/golden/rtrrt_*.npz, from oracle/_ref): node set, parents, cost bits, collision flags,
queue lengths, rand() consumption and the returned path, tick after tick."""
import os

import numpy as np
import pytest

from helpers import TESTD_OUTLINES, dyn_plane, golden_cases

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("case", golden_cases()["rtrrt"], ids=lambda c: c["name"])
def test_cpp_planner_reproduces_reference_golden(oracle, case):
    from rtpp_b200 import capi

    g = np.load(os.path.join(GOLD, f"rtrrt_{case['name']}.npz"))
    S = oracle.load_map(case["map"])
    pl = capi.PlannerRTRRT(S)
    pl.set_clock(case["dt"])
    pl.set_rand(oracle.rand31(case["seed"], case["n_rand"]))
    pl.set_agent(*case["agent"])
    pl.set_goal(*case["goal"])
    rng = np.random.default_rng(case["seed"])
    for tick in range(case["ticks"]):
        if tick in case["dobs_ticks"]:
            pl.set_dobs(dyn_plane(S.shape, rng, n_blobs=6))
        pl.update_tree()
        pl.expand_rewiring(50)
        path, col, cost = pl.get_path()
        xy, par, c, cl = pl.dump()
        k = f"t{tick}_"
        assert pl.rand_pos() == int(g[k + "rand_pos"])
        assert np.array_equal(xy, g[k + "xy"]) and np.array_equal(par, g[k + "par"])
        assert np.array_equal(c.view(np.uint32), g[k + "cost"].view(np.uint32)) and np.array_equal(cl, g[k + "col"])
        assert tuple(g[k + "queues"]) == pl.queue_sizes()
        assert bool(g[k + "path_col"]) == col and np.array_equal(path, g[k + "path"])
        assert np.float32(cost).view(np.uint32) == g[k + "path_cost"].view(np.uint32)
        if tick in case["move_ticks"]:
            t = xy[(len(xy) * 2) // 3]
            pl.set_agent(float(t[0]), float(t[1]))
    pl.close()


def test_cpp_planner_main_task_with_obstacles(oracle):
    """MainTask() = DrawDObsmap + UpdateTree + ExpandRewiring(50) with the scripted GAME0 obstacle
    set; the drawn plane equals the oracle rasteriser's and single-edge CrashDetection agrees."""
    from rtpp_b200 import capi

    S = oracle.load_map("GAME0")
    pl = capi.PlannerRTRRT(S)
    pl.set_clock(20000)
    pl.set_rand(oracle.rand31(5, 1_000_000))
    pl.set_agent(350.0, 2350.0)
    pl.set_goal(1400.0, 2050.0)
    obs = [(0, 370, 388, 0.0, 57, [(0, 0)]), (0, 2370, 630, 0.0, 114, [(0, 0)]),
           (1, 1364, 1364, 0.78539816, 0, TESTD_OUTLINES[3]), (1, 1340, 2520, 0.0, 0, TESTD_OUTLINES[4])]
    want = np.zeros(S.shape, np.uint8)
    for num, (typ, x, y, ang, rad, outline) in enumerate(obs):
        pl.add_obstacle(num, typ, x, y, ang, rad, outline)
        c, s = np.float32(np.cos(np.float32(ang))), np.float32(np.sin(np.float32(ang)))
        o = np.array(outline, np.float32)
        px = (c * o[:, 0] - s * o[:, 1] + np.float32(x)).astype(np.int32)
        py = (s * o[:, 0] + c * o[:, 1] + np.float32(y)).astype(np.int32)
        if typ == 0:
            for qx, qy in zip(px, py):
                oracle.raster_circle(want, int(qx), int(qy), int(rad))
        else:
            oracle.raster_fillpoly(want, np.stack([px, py], 1))
    for _ in range(3):
        pl.main_task()
    assert np.array_equal(pl.get_dobs(), want)
    assert pl.num_nodes() > 20
    xy = pl.dump()[0]
    e = np.concatenate([xy[:-1], xy[1:]], 1)[:200]
    want_v = oracle.edges_check(S, want, e, 5, 3)
    got_v = np.array([pl.crash(r[:2], r[2:]) for r in e], np.uint8)
    assert np.array_equal(got_v, want_v)
    pl.close()


@pytest.mark.parametrize("case", golden_cases()["rtfmt"], ids=lambda c: c["name"])
def test_cpp_rtfmt_planner_reproduces_reference_golden(oracle, case):
    """RTFMTplanner (C++ drop-in): SetAgent/SetGoal/Start with the injected rand stream, then the
    reference's call sequence UpdateTree + 500 x (ExpandFMT, RewireFromObstacles, RewireFromRoot)."""
    from rtpp_b200 import capi

    g = np.load(os.path.join(GOLD, f"rtfmt_{case['name']}.npz"))
    S = oracle.load_map(case["map"])
    pl = capi.PlannerRTFMT(S)
    pl.set_rand(oracle.rand31(case["seed"], case["n_rand"]))
    pl.set_agent(*case["agent"])
    pl.set_goal(*case["goal"])
    pl.start(case["samples"])
    assert pl.rand_pos() == int(g["rand_used"])
    rng = np.random.default_rng(case["seed"])
    for rnd in range(case["rounds"]):
        if rnd in case["dobs_rounds"]:
            pl.set_dobs(dyn_plane(S.shape, rng, n_blobs=5))
        pl.update_tree()
        if rnd == 1:
            pl.expand(2, True)      # method-by-method calls, as a caller of the reference class would
            pl.expand(498, True)
        else:
            pl.expand(500, True)
        xy, par, c, cl, st = pl.dump()
        k = f"r{rnd}_"
        assert np.array_equal(xy, g["xy"]) and np.array_equal(par, g[k + "par"])
        assert np.array_equal(c.view(np.uint32), g[k + "cost"].view(np.uint32))
        assert np.array_equal(cl, g[k + "col"]) and np.array_equal(st, g[k + "state"])
        assert np.array_equal(pl.z(), g[k + "z"])
        p = pl.get_path()
        assert np.array_equal(p, g[k + "path"])
        if rnd in case["move_rounds"] and len(p) > 2:
            pl.set_agent(float(p[1][0]), float(p[1][1]))
    sizes = pl.set_sizes()
    assert sizes["node_map"] == len(xy) and sizes["unv"] + sizes["open"] + sizes["opennew"] + sizes["closed"] == len(xy)
    pl.close()


@pytest.mark.parametrize("case", golden_cases()["amrrt"], ids=lambda c: c["name"])
def test_cpp_amrrt_planner_reproduces_reference_golden(oracle, case):
    """AMRRTplanner (C++ drop-in): MainTask + GetPath, 25 ticks, against the reference's golden."""
    from rtpp_b200 import capi

    g = np.load(os.path.join(GOLD, f"amrrt_{case['name']}.npz"))
    S = oracle.load_map(case["map"])
    pl = capi.PlannerAMRRT(S)
    cells, mat, ratio = case["diffusion"](S.shape[1])
    pl.set_diffusion(cells, mat, ratio)
    pl.set_rand(oracle.rand31(case["seed"], case["n_rand"]))
    pl.set_clock(case["dt"])
    pl.set_agent(*case["agent"])
    pl.set_goal(*case["goal"])
    o = oracle.OrcAMRRT(S)
    o.set_diffusion(cells, mat, ratio)
    assert pl.diffusivity((100, 200), (700, 900)) == o.diffusivity((100, 200), (700, 900))
    rng = np.random.default_rng(case["seed"])
    pl.set_dobs(None)
    for tick in range(case["ticks"]):
        if tick in case["dobs_ticks"]:
            pl.set_dobs(dyn_plane(S.shape, rng, n_blobs=4))
        pl.main_task()
        xy, par, c, cl = pl.dump()
        path, pcol, pcost = pl.get_path()
        k = f"t{tick}_"
        assert pl.rand_pos() == int(g[k + "rand_pos"]), f"tick {tick}"
        assert np.array_equal(xy, g[k + "xy"]) and np.array_equal(par, g[k + "par"])
        assert np.array_equal(c.view(np.uint32), g[k + "cost"].view(np.uint32)) and np.array_equal(cl, g[k + "col"])
        assert np.array_equal(path, g[k + "path"]) and pcol == bool(g[k + "path_col"])
        assert np.float32(pcost).view(np.uint32) == g[k + "path_cost"].view(np.uint32)
    pl.close()


def test_cpp_bimap_reproduces_reference_golden(oracle):
    """The C++ BImap drop-in (planners/CDRT.h) through the reference's own method sequence — DrawBImap, RTRT_AddDynamicObstacle,
    RTRT_SetAgent/SetGoal, RTRT_start, then RTRT_task + GetPath per frame with obstacles mutated in place and the agent
    following its path — against the golden vectors the reference's CDRT.cpp produced (tests/golden/cdrt_mess6_g0.npz)."""
    import cdrt_scenarios as sc
    from rtpp_b200 import capi

    case = sc.GOLDEN_CASES[0]
    z = np.load(os.path.join(GOLD, f"cdrt_{case['name']}.npz"))
    fx = oracle.load_cdrt_fixture(case["map"])
    pl = capi.PlannerCDRT(fx)
    amap, imap = pl.draw_bimap()  # host fillConvexPoly == the real cv2's planes
    assert np.array_equal(amap, fx["amap"]) and np.array_equal(imap, fx["imap"])
    sc.setup(pl, case["map"], case["goal"], oracle.rand31(case["seed"], case["n_rand"]), dt_ns=case["dt"], params=case["params"])
    for k in range(case["ticks"]):
        path, col, cost = sc.tick(pl, case["map"], k)
        st = pl.state()
        assert st["status"] == 0
        r = z["rows"][k]
        assert (st["queue"], st["rand_pos"], st["clock_reads"], st["island"], st["n_trees"], st["samples_left"], int(col), len(path)) == \
               tuple(int(v) for v in (r[1], r[2], r[3], r[4], r[5], r[6], r[9], r[10])), f"tick {k}"
        assert np.array_equal(path, z[f"t{k}_path"]) and np.float64(cost).tobytes() == z[f"t{k}_cost"].tobytes(), f"tick {k} path"
        if k in case["dump_ticks"]:
            for i, a in enumerate(pl.dump()):  # Node_Map with parents, costs, invM and subs sizes
                assert np.array_equal(a.view(np.uint8), z[f"t{k}_dump{i}"].view(np.uint8)), f"tick {k} dump {i}"
    pl.close()


def test_cpp_planner_device_drawn_plane_equals_uploaded_plane(oracle):
    """DrawDObsmap() ships the obstacle shapes and the library draws the ctx and engine planes on the device; a second
    planner gets the SAME plane as uploaded bytes (SetDObsmap of the first one's public DObsmap member).  Both must
    grow identical trees over moving obstacles — the device rasteriser, the host rasteriser and the upload path agree
    wherever an edge probes them."""
    from rtpp_b200 import capi

    S = oracle.load_map("GAME0")
    rand = oracle.rand31(17, 1_500_000)
    pls = [capi.PlannerRTRRT(S), capi.PlannerRTRRT(S)]
    obs = [(0, 700, 2300, 0.0, 90, [(0, 0), (40, 10)]), (1, 1000, 2200, 0.3, 0, TESTD_OUTLINES[1]), (1, 1364, 1364, 0.78539816, 0, TESTD_OUTLINES[3])]
    for pl in pls:
        pl.set_clock(20000)
        pl.set_rand(rand)
        pl.set_agent(350.0, 2350.0)
        pl.set_goal(1400.0, 2050.0)
    for num, (typ, x, y, ang, rad, outline) in enumerate(obs):
        pls[0].add_obstacle(num, typ, x, y, ang, rad, outline)
    for tick in range(6):
        for num, (typ, x, y, ang, rad, outline) in enumerate(obs):  # the obstacles move and turn
            pls[0].add_obstacle(num, typ, x + 15 * tick, y - 10 * tick, ang + 0.2 * tick, rad, outline)
        pls[0].main_task()
        plane = pls[0].get_dobs()
        assert plane.any()
        pls[1].set_dobs(plane)
        pls[1].update_tree()
        pls[1].expand_rewiring(50)
        a, b = pls[0].dump(), pls[1].dump()
        assert all(np.array_equal(u.view(np.uint8), v.view(np.uint8)) for u, v in zip(a, b)), tick
        assert pls[0].rand_pos() == pls[1].rand_pos()
        assert pls[0].get_path()[2] == pls[1].get_path()[2]
    assert len(a[0]) > 50 and a[3].any()  # some nodes are currently behind an obstacle
    for pl in pls:
        pl.close()

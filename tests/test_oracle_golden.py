"""CPU: the oracle restatement reproduces golden vectors that were produced by the REFERENCE'S OWN
sources (compiled in place into oracle/_ref by oracle/Makefile; generator: tools/make_golden.py).
This is the pin of the oracle for the in-tree arithmetic: edge checks (E1-E4), distance (C2), cost
algebra (C1) and the RT-RRT*, RT-FMT and AM-RRT* loops (R1-R3, F1-F2, P1)."""
import os

import numpy as np
import pytest

from helpers import dyn_plane, edge_mix, golden_cases

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _bits(a, n):
    return np.unpackbits(a)[:n]


@pytest.mark.parametrize("name", ["GAME0", "MESS6", "FLOOR5", "MAZE0"])
def test_edge_checks_match_reference(oracle, name):
    g = np.load(os.path.join(GOLD, f"edges_{name}.npz"))
    S = oracle.load_map(name)
    H, W = S.shape
    D = dyn_plane(S.shape, np.random.default_rng(7))
    e = edge_mix(np.random.default_rng(123), 200_000, W, H)
    n = len(e)
    assert np.array_equal(oracle.edges_check(S, D, e, 5, 3, nthreads=4), _bits(g["e1"], n))   # RTRRT.cpp:202 / RTFMT.cpp:242
    assert np.array_equal(oracle.edges_check(S, D, e, 5, 1, nthreads=4), _bits(g["e2"], n))   # RTFMT.cpp:264
    assert np.array_equal(oracle.edges_check(S, D, e, 2, 3, nthreads=4), _bits(g["e3"], n))   # AMRRT.cpp:266
    m = len(g["e4_pstart"])
    b, p = oracle.edges_first_hit(S, D, e[:m])                                                 # AMRRT.cpp:287
    assert np.array_equal(b, _bits(g["e4_blocked"], m))
    assert np.array_equal(p.view(np.uint32), g["e4_pstart"].view(np.uint32))
    assert np.array_equal(oracle.dist(e[:m]).view(np.uint32), g["dist"].view(np.uint32))      # RTRRT.h:47


@pytest.mark.parametrize("case", golden_cases()["rtrrt"], ids=lambda c: c["name"])
def test_rtrrt_matches_reference(oracle, case):
    g = np.load(os.path.join(GOLD, f"rtrrt_{case['name']}.npz"))
    S = oracle.load_map(case["map"])
    orc = oracle.OrcRTRRT(S)
    orc.set_clock(case["dt"])
    r = oracle.rand31(case["seed"], case["n_rand"])
    assert orc.set_agent(*case["agent"]) and orc.set_goal(*case["goal"])
    rng = np.random.default_rng(case["seed"])
    pos = 0
    for tick in range(case["ticks"]):
        if tick in case["dobs_ticks"]:
            orc.set_dobs(dyn_plane(S.shape, rng, n_blobs=6))
        orc.update_tree()
        pts, offs = orc.presample(r[pos:], 6000)
        used = orc.expand_rewiring(50, pts)
        pos += int(offs[used])
        path, col, cost, cb = orc.get_path()
        xy, par, c, cl = orc.dump()
        k = f"t{tick}_"
        assert pos == int(g[k + "rand_pos"])
        assert np.array_equal(xy, g[k + "xy"]) and np.array_equal(par, g[k + "par"])
        assert np.array_equal(c.view(np.uint32), g[k + "cost"].view(np.uint32)) and np.array_equal(cl, g[k + "col"])
        assert tuple(g[k + "queues"]) == orc.queue_sizes()
        assert bool(g[k + "path_col"]) == col and np.array_equal(path, g[k + "path"])
        assert np.float32(cost).view(np.uint32) == g[k + "path_cost"].view(np.uint32)
        assert np.float32(cb).view(np.uint32) == g[k + "c_base"].view(np.uint32)
        if tick in case["move_ticks"]:
            t = xy[(len(xy) * 2) // 3]
            orc.set_agent(float(t[0]), float(t[1]))


@pytest.mark.parametrize("case", golden_cases()["rtfmt"], ids=lambda c: c["name"])
def test_rtfmt_matches_reference(oracle, case):
    g = np.load(os.path.join(GOLD, f"rtfmt_{case['name']}.npz"))
    S = oracle.load_map(case["map"])
    orc = oracle.OrcRTFMT(S)
    r = oracle.rand31(case["seed"], case["n_rand"])
    assert orc.set_agent(*case["agent"]) and orc.set_goal(*case["goal"])
    assert orc.start(r, case["samples"]) == int(g["rand_used"])
    rng = np.random.default_rng(case["seed"])
    for rnd in range(case["rounds"]):
        if rnd in case["dobs_rounds"]:
            orc.set_dobs(dyn_plane(S.shape, rng, n_blobs=5))
        orc.update_tree()
        orc.expand(500, True)
        xy, par, c, cl, st = orc.dump()
        k = f"r{rnd}_"
        assert np.array_equal(xy, g["xy"]) and np.array_equal(par, g[k + "par"])
        assert np.array_equal(c.view(np.uint32), g[k + "cost"].view(np.uint32))
        assert np.array_equal(cl, g[k + "col"]) and np.array_equal(st, g[k + "state"])
        assert np.array_equal(orc.z(), g[k + "z"]) and np.array_equal(orc.get_path(), g[k + "path"])
        if rnd in case["move_rounds"]:
            p = orc.get_path()
            if len(p) > 2:
                orc.set_agent(float(p[1][0]), float(p[1][1]))


@pytest.mark.parametrize("case", golden_cases()["amrrt"], ids=lambda c: c["name"])
def test_amrrt_matches_reference(oracle, case):
    g = np.load(os.path.join(GOLD, f"amrrt_{case['name']}.npz"))
    S = oracle.load_map(case["map"])
    orc = oracle.OrcAMRRT(S)
    cells, mat, ratio = case["diffusion"](S.shape[1])
    orc.set_diffusion(cells, mat, ratio)
    orc.set_rand(oracle.rand31(case["seed"], case["n_rand"]))
    orc.set_clock(case["dt"])
    assert orc.set_agent(*case["agent"]) and orc.set_goal(*case["goal"])
    rng = np.random.default_rng(case["seed"])
    for tick in range(case["ticks"]):
        if tick in case["dobs_ticks"]:
            orc.set_dobs(dyn_plane(S.shape, rng, n_blobs=4))
        orc.main_task()
        xy, par, c, cl = orc.dump()
        path, pcol, pcost = orc.get_path()
        k = f"t{tick}_"
        assert orc.rand_pos() == int(g[k + "rand_pos"])
        assert np.array_equal(xy, g[k + "xy"]) and np.array_equal(par, g[k + "par"])
        assert np.array_equal(c.view(np.uint32), g[k + "cost"].view(np.uint32)) and np.array_equal(cl, g[k + "col"])
        assert np.array_equal(path, g[k + "path"]) and pcol == bool(g[k + "path_col"])
        assert np.float32(pcost).view(np.uint32) == g[k + "path_cost"].view(np.uint32)

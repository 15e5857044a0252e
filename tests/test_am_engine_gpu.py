"""AM-RRT* engine parity: the GPU engine (Expand with Steer + diffusion metric, RewireRoot, UpdateTree,
GetPath; rand() consumed on the device) against the golden vectors produced by the REFERENCE's own
AMRRT.cpp (tests/golden/amrrt_*.npz): node set, parents, cost bits, flags, rand() position, path."""
import os

import numpy as np
import pytest

from helpers import dyn_plane, golden_cases

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("case", golden_cases()["amrrt"], ids=lambda c: c["name"])
def test_am_engine_reproduces_reference_golden(ctx, oracle, case):
    import rtpp_b200

    g = np.load(os.path.join(GOLD, f"amrrt_{case['name']}.npz"))
    S = oracle.load_map(case["map"])
    H, W = S.shape
    cap_rand = 200_000
    eng = rtpp_b200.Engine(ctx, 1, 8192, 5, W, H, cap_rand, kind=2)
    eng.set_map(0, 1, S)
    cells, mat, ratio = case["diffusion"](W)
    eng.set_diffusion(cells, mat, ratio)
    eng.init_instances(np.zeros(1, np.int32), np.array([case["agent"]], np.float32), np.array([case["goal"]], np.float32))
    r = oracle.rand31(case["seed"], case["n_rand"])
    pos = 0
    rng = np.random.default_rng(case["seed"])
    for tick in range(case["ticks"]):
        if tick in case["dobs_ticks"]:
            eng.set_map(0, 2, dyn_plane(S.shape, rng, n_blobs=4))
        chunk = np.zeros((1, cap_rand), np.int32)
        n = min(cap_rand, len(r) - pos)
        chunk[0, :n] = r[pos: pos + n]
        eng.set_am_inputs(rand=chunk, n_rand=np.array([n], np.int32), outpath_nonempty=np.zeros(1, np.int32))
        eng.tick(update_tree=True, run_ms=80, dt_ns=case["dt"], get_path=True)
        d = dict(zip(eng.STAT, eng.stats()[0].tolist()))
        assert d["status"] == 0 and d["rand_exhausted"] == 0
        pos += d["rand_pos"]
        xy, par, c, cl = eng.read_tree(0)
        k = f"t{tick}_"
        assert pos == int(g[k + "rand_pos"]), f"tick {tick}"
        assert np.array_equal(xy, g[k + "xy"]) and np.array_equal(par, g[k + "par"])
        assert np.array_equal(c.view(np.uint32), g[k + "cost"].view(np.uint32)) and np.array_equal(cl, g[k + "col"])
        assert bool(d["path_col"]) == bool(g[k + "path_col"])
        if not d["path_col"]:
            ids = eng.read_path(0)
            p = np.array([case["goal"] if j == -2 else xy[j] for j in ids], np.float32)
            assert np.array_equal(p, g[k + "path"])
            assert np.uint32(d["path_cost_bits"]) == g[k + "path_cost"].view(np.uint32)
    assert len(xy) > (80 if case["name"] == "mess6" else 60)  # the tree really grew
    eng.close()


def _ellipse(c_base, goal, root):
    """Host side of the informed sampler (AMRRT.cpp:418-432) with glibc float functions."""
    import ctypes as C

    m = C.CDLL("libm.so.6")
    for f in (m.cosf, m.sinf, m.sqrtf):
        f.restype, f.argtypes = C.c_float, [C.c_float]
    m.atan2f.restype, m.atan2f.argtypes = C.c_float, [C.c_float, C.c_float]
    f32 = np.float32
    dx, dy = f32(goal[0]) - f32(root[0]), f32(goal[1]) - f32(root[1])
    c_min = f32(m.sqrtf(f32(f32(dx * dx) + f32(dy * dy))))
    a = f32(f32(c_base) / f32(2.0))
    with np.errstate(invalid="ignore"):
        b = f32(f32(m.sqrtf(f32(f32(f32(c_base) * f32(c_base)) - f32(c_min * c_min)))) / f32(2.0))
    ang = f32(m.atan2f(dy, dx))
    return [a, b, f32(m.cosf(ang)), f32(m.sinf(ang)), f32(f32(f32(goal[0]) + f32(root[0])) / f32(2)), f32(f32(f32(goal[1]) + f32(root[1])) / f32(2))]


def test_am_engine_informed_sampling_against_oracle(ctx, oracle):
    """OutPath filled by the caller (AMRRT/test/testD.cpp) switches the sampler to the informed ellipse: the engine evaluates the
    ellipse (atan2f, cosf, sinf of the axis) and the cosf / sinf of every drawn angle on the device with glibc's bits
    (csrc/glibc_math.cuh) — even ticks; odd ticks take the ellipse from the caller (host libm) — against the oracle."""
    import rtpp_b200
    from helpers import synthetic_diffusion

    S = oracle.load_map("MESS6")
    H, W = S.shape
    cap_rand = 200_000
    agent, goal = (720.0, 720.0), (1240.0, 1240.0)
    eng = rtpp_b200.Engine(ctx, 1, 8192, 5, W, H, cap_rand, kind=2)
    eng.set_map(0, 1, S)
    cells, mat, ratio = synthetic_diffusion(W)
    eng.set_diffusion(cells, mat, ratio)
    eng.init_instances(np.zeros(1, np.int32), np.array([agent], np.float32), np.array([goal], np.float32))
    orc = oracle.OrcAMRRT(S)
    orc.set_diffusion(cells, mat, ratio)
    r = oracle.rand31(33, 3_000_000)
    orc.set_rand(r)
    orc.set_clock(500000, 0)
    orc.set_agent(*agent)
    orc.set_goal(*goal)
    pos, informed, found, c_base = 0, 0, False, -1.0
    for tick in range(30):
        chunk = np.zeros((1, cap_rand), np.int32)
        chunk[0] = r[pos: pos + cap_rand]
        eng.set_am_inputs(rand=chunk, n_rand=np.array([cap_rand], np.int32), outpath_nonempty=np.array([int(found)], np.int32))
        if tick % 2:
            eng.tick(update_tree=True, run_ms=0, dt_ns=500000, get_path=False)  # UpdateTree first: the ellipse needs the root
            root = eng.read_tree(0)[0][eng.stats()[0][1]]
            eng.set_am_inputs(ellipse=np.array([_ellipse(c_base, goal, root)], np.float32))
            eng.tick(update_tree=False, run_ms=80, dt_ns=500000, get_path=True)
        else:
            eng.tick(update_tree=True, run_ms=80, dt_ns=500000, get_path=True)
        orc.main_task()
        d = dict(zip(eng.STAT, eng.stats()[0].tolist()))
        pos += d["rand_pos"]
        assert d["status"] == 0 and pos == orc.rand_pos(), f"tick {tick}"
        for x, y in zip(orc.dump(), eng.read_tree(0)):
            assert np.array_equal(x.view(np.uint8), y.view(np.uint8)), f"tick {tick}"
        opath, ocol, ocost = orc.get_path()
        assert bool(d["path_col"]) == ocol
        c_base = orc.c_base()
        assert np.float32(c_base).view(np.uint32) == np.uint32(d["c_base_bits"])
        found = not ocol
        orc.set_outpath(found)
        informed += int(found)
    assert informed >= 5
    eng.close()


def test_am_engine_rewire_goal_when_the_goal_is_a_tree_node(ctx, oracle):
    """RewireGoal (AMRRT.cpp:521-585) only runs while the goal point itself is a collision-free tree node, which Expand
    never produces but a caller can: after a few ticks the goal is moved onto an existing node.  Engine vs oracle,
    tick by tick (trees, rand() position), and the goal stack / queue must have been in use."""
    import rtpp_b200
    from helpers import synthetic_diffusion

    S = oracle.load_map("MESS6")
    H, W = S.shape
    cap_rand = 200_000
    agent, goal = (720.0, 720.0), (1240.0, 1240.0)
    eng = rtpp_b200.Engine(ctx, 1, 8192, 5, W, H, cap_rand, kind=2)
    eng.set_map(0, 1, S)
    cells, mat, ratio = synthetic_diffusion(W)
    eng.set_diffusion(cells, mat, ratio)
    eng.init_instances(np.zeros(1, np.int32), np.array([agent], np.float32), np.array([goal], np.float32))
    orc = oracle.OrcAMRRT(S)
    orc.set_diffusion(cells, mat, ratio)
    r = oracle.rand31(77, 3_000_000)
    orc.set_rand(r)
    orc.set_clock(500000, 0)
    orc.set_agent(*agent)
    orc.set_goal(*goal)
    pos, used = 0, 0
    for tick in range(16):
        if tick == 5:  # move the goal onto a node far from the root
            xy, _, _, col = eng.read_tree(0)
            d2 = ((xy - np.array(agent, np.float32)) ** 2).sum(1)
            j = int(np.argmax(np.where(col == 0, d2, -1.0)))  # a collision-free node far from the root
            orc.set_goal(float(xy[j, 0]), float(xy[j, 1]))
            eng.set_inputs(goal_xy=xy[j: j + 1])
        chunk = np.zeros((1, cap_rand), np.int32)
        chunk[0] = r[pos: pos + cap_rand]
        eng.set_am_inputs(rand=chunk, n_rand=np.array([cap_rand], np.int32), outpath_nonempty=np.zeros(1, np.int32))
        eng.tick(update_tree=True, run_ms=80, dt_ns=500000, get_path=True)
        orc.main_task()
        opath, ocol, ocost = orc.get_path()
        d = dict(zip(eng.STAT, eng.stats()[0].tolist()))
        pos += d["rand_pos"]
        assert d["status"] == 0 and pos == orc.rand_pos(), f"tick {tick}"
        for x, y in zip(orc.dump(), eng.read_tree(0)):
            assert np.array_equal(x.view(np.uint8), y.view(np.uint8)), f"tick {tick}"
        assert bool(d["path_col"]) == ocol and np.float32(orc.c_base()).view(np.uint32) == np.uint32(d["c_base_bits"])
        used = d["pops_r"]
    assert used > 100  # nodes popped from Sgoal / Qgoal: RewireGoal ran
    eng.close()

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
    assert len(xy) > 80
    eng.close()

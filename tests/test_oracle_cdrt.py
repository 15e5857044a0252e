"""CPU: the CDRT-RRT* oracle (oracle/cdrt_oracle.hpp) against (a) golden vectors produced by the reference's own
CDRT.cpp compiled in place (tools/make_cdrt_golden.py), (b) that compiled reference live where oracle/_ref
exists, and (c) the real cv2.fillConvexPoly id planes committed as fixtures (tools/make_cdrt_fixtures.py)."""
import ctypes as C
import os

import numpy as np
import pytest

import cdrt_scenarios as sc

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libref_cdrt.so")
needs_ref = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref not built")


def check_against_golden(pl, case, z, rand_pos, clock_reads):
    sxy, scut = pl.samples()
    assert np.array_equal(sxy, z["samples_xy"]) and np.array_equal(scut, z["samples_cut"])
    for i, a in enumerate(pl.dump()):
        assert np.array_equal(a.view(np.uint8), z[f"start_dump{i}"].view(np.uint8)), f"start dump {i}"
    for k in range(case["ticks"]):
        path, col, cost = sc.tick(pl, case["map"], k)
        d, st = pl.dump(), pl.state()
        row = z["rows"][k]
        got = (len(d[0]), st["queue"], rand_pos(), clock_reads(), st["island"], st["n_trees"], st["samples_left"], st["larger"],
               st["smaller"], int(col), len(path))
        assert got == tuple(int(v) for v in row), f"tick {k}: {got} vs {tuple(row)}"
        assert np.array_equal(path, z[f"t{k}_path"]), f"tick {k} path"
        assert np.float64(cost).tobytes() == z[f"t{k}_cost"].tobytes(), f"tick {k} path cost"
        if k in case["dump_ticks"]:
            for i, a in enumerate(d):
                assert np.array_equal(a.view(np.uint8), z[f"t{k}_dump{i}"].view(np.uint8)), f"tick {k} dump {i}"
            isl, cnt = pl.tree_islands()
            assert np.array_equal(isl, z[f"t{k}_islands"]) and np.array_equal(cnt, z[f"t{k}_island_pts"])
        assert np.array_equal(sc.dump_hash(d), z[f"t{k}_hash"]), f"tick {k} node dump hash"


@pytest.mark.parametrize("case", sc.GOLDEN_CASES, ids=lambda c: c["name"])
def test_oracle_reproduces_reference_golden(oracle, case):
    z = np.load(os.path.join(HERE, "golden", f"cdrt_{case['name']}.npz"))
    fx = oracle.load_cdrt_fixture(case["map"])
    o = oracle.CdrtCpu(fx, "oracle")
    sc.setup(o, case["map"], case["goal"], oracle.rand31(case["seed"], case["n_rand"]), dt_ns=case["dt"], params=case["params"])
    check_against_golden(o, case, z, o.rand_pos, o.clock_reads)
    assert o.error() == 0
    o.close()


@pytest.mark.parametrize("name", ["MESS6", "FLOOR5", "GAME0", "MAZE0"])
def test_fill_convex_poly_matches_cv2_planes(oracle, name):
    """DrawBImap (CDRT.cpp:121-175) through the oracle rasteriser == the planes the real cv2.fillConvexPoly drew."""
    fx = oracle.load_cdrt_fixture(name)
    H, W = fx["H"], fx["W"]
    amap = np.full((H, W), 0xFFFF, np.uint16)
    imap = np.full((H, W), 0xFFFF, np.uint16)
    for gi, g in enumerate(fx["graphs"]):
        for pi in range(len(g["poly_core"])):
            v = np.ascontiguousarray(g["pv"][g["pv_off"][pi]:g["pv_off"][pi + 1]].astype(np.int32))
            for plane, val in ((amap, gi), (imap, pi)):
                oracle.lib().orc_fill_convex16(plane.ctypes.data_as(C.c_void_p), H, W, v.ctypes.data_as(C.c_void_p), len(v), val)
    assert np.array_equal(amap, fx["amap"]) and np.array_equal(imap, fx["imap"])


@needs_ref
@pytest.mark.parametrize("name,goal,seed,dt,ticks", [("GAME0", 2, 3, 5000, 25), ("MAZE0", 1, 5, 20000, 20)])
def test_oracle_vs_compiled_reference_live(oracle, name, goal, seed, dt, ticks):
    fx = oracle.load_cdrt_fixture(name)
    r = oracle.rand31(seed, 600_000)
    ref, o = oracle.CdrtCpu(fx, "ref"), oracle.CdrtCpu(fx, "oracle")
    for pl in (ref, o):
        sc.setup(pl, name, goal, r, dt_ns=dt)
    rng = np.random.default_rng(seed)
    for k in range(ticks):
        pa, pb = sc.tick(ref, name, k), sc.tick(o, name, k)
        assert np.array_equal(pa[0], pb[0]) and pa[1] == pb[1] and np.float64(pa[2]).tobytes() == np.float64(pb[2]).tobytes()
        da, db = ref.dump(), o.dump()
        assert len(da[0]) == len(db[0])
        for x, y in zip(da, db):
            assert np.array_equal(x.view(np.uint8), y.view(np.uint8)), f"tick {k}"
        sa, sb = ref.state(), o.state()
        assert all(np.array_equal(sa[key], sb[key]) for key in sa), f"tick {k}: {sa} vs {sb}"
        assert ref.rand_pos() == o.rand_pos() and ref.clock_reads() == o.clock_reads()
        if k % 5 == 4:
            # E5 on its own: the full RTRT_CrashDetection (topology switch + fp64 march) on random node pairs,
            # and RTRT_Neighbor for random nodes, in the planner's current state
            xy, _, _, _, inv, _ = da
            i, j = rng.integers(0, len(xy), 4000), rng.integers(0, len(xy), 4000)
            e = np.concatenate([xy[i], xy[j]], axis=1)
            assert np.array_equal(ref.crash(e, inv[i], inv[j]), o.crash(e, inv[i], inv[j]))
            for q in rng.integers(0, len(xy), 40):
                na, nb = ref.neighbor(xy[q, 0], xy[q, 1], int(inv[q])), o.neighbor(xy[q, 0], xy[q, 1], int(inv[q]))
                assert np.array_equal(na, nb)
    assert np.array_equal(ref.draw_dobs(), o.draw_dobs())
    for n in range(4):
        assert np.array_equal(ref.check_points(n), o.check_points(n))
    ref.close()
    o.close()

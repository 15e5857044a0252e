"""CPU: live comparison of the oracle restatement with oracle/_ref (the reference's own planner
sources compiled in place).  Skipped where oracle/_ref has not been built (it needs /root/reference
at build time; the prebuilt library travels to the GPU box)."""
import os

import numpy as np
import pytest

from helpers import dyn_plane, edge_mix

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(os.path.dirname(HERE), "oracle", "_ref")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "libref_rtrrt.so")), reason="oracle/_ref not built")


@pytest.mark.parametrize("name", ["FLOOR5", "Trap", "dilated"])
def test_edge_routines_live(oracle, name):
    S = oracle.load_map(name)
    H, W = S.shape
    rng = np.random.default_rng(len(name))
    D = dyn_plane(S.shape, rng, n_blobs=5)
    e = edge_mix(rng, 150_000, W, H, short_px=min(80.0, W / 8))
    rt, fm, am = oracle.RefRTRRT(S), oracle.RefRTFMT(S), oracle.RefAMRRT(S)
    for p in (rt, fm, am):
        p.set_dobs(D)
    assert np.array_equal(rt.crash(e), oracle.edges_check(S, D, e, 5, 3))
    assert np.array_equal(fm.crash(e, ndobs=True), oracle.edges_check(S, D, e, 5, 1))
    assert np.array_equal(am.crash(e), oracle.edges_check(S, D, e, 2, 3))
    rb, rp = am.crash_start(e[:30_000])
    ob, op = oracle.edges_first_hit(S, D, e[:30_000])
    assert np.array_equal(rb, ob) and np.array_equal(rp.view(np.uint32), op.view(np.uint32))
    # the algorithmic probe count formula against the reference's loop bounds
    assert (oracle.edges_nprobes(e[:1000], 5) >= 3).all()
    for p in (rt, fm, am):
        p.close()


def test_rtrrt_live_with_informed_sampling(oracle):
    """Longer RT-RRT* run on FLOOR5 where a path is found, so the informed ellipse sampler
    (RTRRT.cpp:243-261, glibc sinf/cosf/atan2) is exercised against the reference."""
    S = oracle.load_map("FLOOR5")
    ref, orc = oracle.RefRTRRT(S), oracle.OrcRTRRT(S)
    r = oracle.rand31(21, 4_000_000)
    ref.set_rand(r)
    ref.set_clock(2000, 0)
    orc.set_clock(2000, 0)
    for p in (ref, orc):
        p.set_agent(1000.0, 1200.0)
        p.set_goal(1750.0, 900.0)
    pos, informed = 0, 0
    for tick in range(24):
        ref.update_tree()
        orc.update_tree()
        pts, offs = orc.presample(r[pos:], 8000)
        ref.expand_rewiring(50)
        used = orc.expand_rewiring(50, pts)
        pos += int(offs[used])
        assert pos == ref.rand_pos()
        rp, op = ref.get_path(), orc.get_path()
        assert np.array_equal(rp[0], op[0]) and rp[1:] == op[1:]
        informed += int(op[3] >= 0)
        for x, y in zip(ref.dump(), orc.dump()):
            assert np.array_equal(x.view(np.uint8), y.view(np.uint8))
    assert informed >= 2, "the scenario must reach the informed-sampling branch"


def test_amrrt_live_with_informed_sampling(oracle):
    """AM-RRT* with OutPath filled by the caller (as AMRRT/test/testD.cpp does): the informed-ellipse
    sampler branch (AMRRT.cpp:415-441) of the oracle against the compiled reference."""
    from helpers import synthetic_diffusion

    S = oracle.load_map("MESS6")
    ref, orc = oracle.RefAMRRT(S), oracle.OrcAMRRT(S)
    cells, mat, ratio = synthetic_diffusion(S.shape[1])
    r = oracle.rand31(33, 3_000_000)
    for p in (ref, orc):
        p.set_diffusion(cells, mat, ratio)
        p.set_rand(r)
        p.set_clock(500000, 0)
        p.set_agent(720.0, 720.0)
        p.set_goal(1240.0, 1240.0)
    informed = 0
    for tick in range(30):
        for p in (ref, orc):
            p.main_task()
        rp, op = ref.get_path(), orc.get_path()
        assert np.array_equal(rp[0], op[0]) and rp[1:] == op[1:]
        found = not op[1]
        for p in (ref, orc):
            p.set_outpath(found)
        informed += int(found)
        assert ref.rand_pos() == orc.rand_pos()
        for x, y in zip(ref.dump(), orc.dump()):
            assert np.array_equal(x.view(np.uint8), y.view(np.uint8))
    assert informed >= 5


def test_amrrt_live_rewire_goal(oracle):
    """RewireGoal (AMRRT.cpp:521-585) only runs while the goal point itself is a collision-free tree node — Expand
    never makes it one, a caller can.  After a few ticks the goal is moved onto an existing node: the oracle's
    RewireGoal against the compiled reference (trees bit for bit, same rand() consumption, same clock reads)."""
    from helpers import synthetic_diffusion

    S = oracle.load_map("MESS6")
    ref, orc = oracle.RefAMRRT(S), oracle.OrcAMRRT(S)
    cells, mat, ratio = synthetic_diffusion(S.shape[1])
    r = oracle.rand31(77, 3_000_000)
    for p in (ref, orc):
        p.set_diffusion(cells, mat, ratio)
        p.set_rand(r)
        p.set_clock(500000, 0)
        p.set_agent(720.0, 720.0)
        p.set_goal(1240.0, 1240.0)
    moved = False
    for tick in range(16):
        if tick == 5:
            xy, _, _, col = orc.dump()
            d2 = ((xy - np.array([720.0, 720.0], np.float32)) ** 2).sum(1)
            j = int(np.argmax(np.where(col == 0, d2, -1.0)))
            for p in (ref, orc):
                p.set_goal(float(xy[j, 0]), float(xy[j, 1]))
            moved = True
        for p in (ref, orc):
            p.main_task()
        rp, op = ref.get_path(), orc.get_path()
        assert np.array_equal(rp[0], op[0]) and rp[1:] == op[1:]
        assert ref.rand_pos() == orc.rand_pos()
        for x, y in zip(ref.dump(), orc.dump()):
            assert np.array_equal(x.view(np.uint8), y.view(np.uint8)), f"tick {tick}"
    assert moved and ref.L.ref_clock_reads() > 0
    # witness that RewireGoal ran and rewired: without the goal on a node the same stream grows a different tree
    ctl = oracle.OrcAMRRT(S)
    ctl.set_diffusion(cells, mat, ratio)
    ctl.set_rand(r)
    ctl.set_clock(500000, 0)
    ctl.set_agent(720.0, 720.0)
    ctl.set_goal(1240.0, 1240.0)
    for tick in range(16):
        ctl.main_task()
        ctl.get_path()
    a, b = orc.dump(), ctl.dump()
    assert len(a[0]) != len(b[0]) or any(not np.array_equal(x.view(np.uint8), y.view(np.uint8)) for x, y in zip(a, b))

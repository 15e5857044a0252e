"""CPU: the oracle's grid-accelerated k-NN equals its brute-force definition, and both agree with
an independent implementation (scipy cKDTree) on neighbour sets where distances are distinct."""
import numpy as np
import pytest


def _pts(rng, n, w, h):
    return (rng.random((n, 2), dtype=np.float32) * np.array([w, h], np.float32)).astype(np.float32)


@pytest.mark.parametrize("n,k", [(1, 25), (10, 25), (25, 25), (500, 1), (5000, 20), (20000, 25)])
def test_grid_equals_brute(oracle, n, k):
    rng = np.random.default_rng(n + k)
    pts = _pts(rng, n, 1000, 700)
    q = np.concatenate([rng.random((300, 2)) * np.array([1000, 700]), pts[:100].astype(np.float64),
                        np.array([[-20.0, -5.0], [1500.0, 900.0]])])
    bi, bd, bc = oracle.knn(pts, q, k)
    gi, gd, gc = oracle.knn(pts, q, k, grid=(1000, 700, 16))
    assert np.array_equal(bc, gc) and np.array_equal(bi, gi)
    assert np.array_equal(bd.view(np.uint64), gd.view(np.uint64))
    assert (bc == min(k, n)).all()


def test_against_scipy_kdtree(oracle):
    from scipy.spatial import cKDTree

    rng = np.random.default_rng(1)
    pts = _pts(rng, 20000, 2729, 2729)
    q = rng.random((2000, 2)) * 2729
    oi, od, _ = oracle.knn(pts, q, 25, grid=(2729, 2729, 32), nthreads=4)
    dd, ii = cKDTree(pts.astype(np.float64)).query(q, k=25)
    # cKDTree returns Euclidean distances; compare neighbour sets and d2 to rounding
    assert np.array_equal(np.sort(oi, 1), np.sort(ii, 1))
    assert np.allclose(np.sqrt(od), dd, rtol=1e-12, atol=0)
    assert (np.diff(od, axis=1) >= 0).all()


def test_tie_rule_is_lowest_index_first(oracle):
    pts = np.array([[10, 10], [12, 10], [8, 10], [10, 12], [10, 8], [10, 10]], np.float32)
    idx, d2, cnt = oracle.knn(pts, np.array([[10.0, 10.0]]), 4)
    assert idx.tolist() == [[0, 5, 1, 2]] and d2[0].tolist() == [0.0, 0.0, 4.0, 4.0]


def _tie_exposure(oracle, pts, q, k):
    """Queries whose answer depends on a tie rule CGAL leaves unspecified (Orthogonal_k_neighbor_search sorts by distance
    only): 'set' = the k-th and (k+1)-th nearest are equally far, so WHICH nodes are returned is the rule's choice;
    'order' = two returned neighbours are equally far, so only their order is."""
    idx, d2, cnt = oracle.knn(pts, q, k + 1)
    full = cnt == k + 1
    set_amb = full & (d2[:, k - 1] == d2[:, k])
    order_amb = (np.diff(d2[:, :k], axis=1) == 0).any(1)
    return int(set_amb.sum()), int((order_amb & ~set_amb).sum())


def test_knn_tie_exposure_is_quantified(oracle):
    """k-NN parity is pinned to the published definition, not to a CGAL run (CGAL needs Boost, absent here), so
    the exposure to CGAL's unspecified tie order is counted instead of assumed away.  Planner workloads — binary32 sample
    points queried with their own coordinates or with rand()-drawn points — have NO set-ambiguous query in 20 000 and a
    handful of order-only ties at most; an integer lattice, the adversarial case the GPU tests also cover, is ambiguous for
    most queries, and there the (d2, id) rule of this repository is what both the oracle and the kernels implement."""
    rng = np.random.default_rng(12)
    pts = _pts(rng, 100_000, 4096, 4096)
    q = np.concatenate([pts[:10_000].astype(np.float64), rng.random((10_000, 2)) * 4096])
    s, o = _tie_exposure(oracle, pts, q, 25)
    assert s == 0 and o <= 5, (s, o)
    # samples on integer pixels (rand() % shape, RTFMT.cpp:401-405): equal integer d2 are common — order ties, and set ties too
    ipts = np.unique(rng.integers(0, 1000, (20_000, 2)), axis=0).astype(np.float32)
    si, oi = _tie_exposure(oracle, ipts, ipts[:2000].astype(np.float64), 25)
    assert si + oi > 500   # this IS the exposed case: RT-FMT's static 25-NN table on integer samples
    xs, ys = np.meshgrid(np.arange(0, 64, dtype=np.float32) * 4, np.arange(0, 64, dtype=np.float32) * 4)
    lat = np.stack([xs.ravel(), ys.ravel()], 1)
    sl, ol = _tie_exposure(oracle, lat, lat[:1000].astype(np.float64), 25)
    assert sl + ol > 900 and sl > 100   # (k = 25 is exactly the 5 x 5 block of an interior lattice point: set ties come from the border)
    print(f"tie exposure: float samples set {s} order {o}; integer samples set {si} order {oi}; lattice set {sl} order {ol}")


def test_kdtree_form_equals_the_definition(oracle):
    """The reference-shaped index of the CPU baseline (kd-tree, sliding-midpoint splits, bucket 10, rebuilt after inserts:
    SURVEY 8(d)) returns exactly the brute-force answer — ids, order and d2 bits — including ties and coincident points."""
    rng = np.random.default_rng(3)
    pts = _pts(rng, 30_000, 2729, 2729)
    q = np.concatenate([pts[:1500].astype(np.float64), rng.random((1500, 2)) * 2729, np.array([[-50.0, -50.0], [5000.0, 10.0]])])
    for k in (1, 25):
        bi, bd, bc = oracle.knn(pts, q, k, nthreads=4)
        ki, kd, kc = oracle.knn_kdtree(pts, q, k, (2729.0, 2729.0))
        assert np.array_equal(bi, ki) and np.array_equal(bd.view(np.uint64), kd.view(np.uint64)) and np.array_equal(bc, kc)
    xs, ys = np.meshgrid(np.arange(0, 48, dtype=np.float32) * 4, np.arange(0, 48, dtype=np.float32) * 4)
    lat = np.stack([xs.ravel(), ys.ravel()], 1)
    lat = np.concatenate([lat, lat[:100]])[rng.permutation(len(lat) + 100)]   # ties everywhere, 100 duplicated points
    bi, bd, _ = oracle.knn(lat, lat[:600].astype(np.float64), 25)
    ki, kd, _ = oracle.knn_kdtree(lat, lat[:600].astype(np.float64), 25, (192.0, 192.0))
    assert np.array_equal(bi, ki) and np.array_equal(bd.view(np.uint64), kd.view(np.uint64))
    same = np.tile(np.array([[5.0, 5.0]], np.float32), (40, 1))
    assert np.array_equal(oracle.knn(same, same[:3].astype(np.float64), 25)[0], oracle.knn_kdtree(same, same[:3].astype(np.float64), 25, (16.0, 16.0))[0])


def test_planner_loop_is_the_same_on_either_index(oracle):
    """RT-RRT* on the grid index and on the kd-tree rebuilt after every insert: the same tree, byte for byte (only the CPU cost
    differs, which is what bench.py reports as cpu_kdtree_rebuild_*)."""
    S = oracle.load_map("MESS6")
    trees = []
    for mode in (0, 1):
        o = oracle.OrcRTRRT(S)
        o.knn_mode(mode)
        o.set_clock(5000)
        assert o.set_agent(720.0, 720.0) and o.set_goal(70.0, 70.0)
        stream, pos = oracle.rand31(11, 400_000), 0
        for _ in range(6):
            o.update_tree()
            pts, offs = o.presample(stream[pos:], 2048)
            pos += int(offs[o.expand_rewiring(50, pts)])
            o.get_path()
        trees.append(o.dump())
        assert (o.knn_mode(-1) > 100) == (mode == 1)
        o.close()
    assert all(np.array_equal(a.view(np.uint8), b.view(np.uint8)) for a, b in zip(*trees))

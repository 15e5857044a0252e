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

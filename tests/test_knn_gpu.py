"""K2/K3 parity: GPU uniform-grid k-NN (through the C ABI) vs the oracle's binary64 brute force.
Neighbour ids, order and squared distances must be bit-exact (ties: ascending node id)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

W, H = 2729, 2729


def _pts(rng, n, w=W, h=H):
    return (rng.random((n, 2), dtype=np.float32) * np.array([w, h], np.float32)).astype(np.float32)


def _check(ctx, oracle, pts, q, k, grid=None):
    gi, gd, gc = ctx.knn(q, k)
    oi, od, oc = oracle.knn(pts, q, k, grid=grid, nthreads=8)
    assert np.array_equal(gc, oc)
    assert np.array_equal(gi, oi)
    assert np.array_equal(gd.view(np.uint64), od.view(np.uint64))


@pytest.mark.parametrize("n", [1, 5, 24, 25, 26, 1000, 30000])
@pytest.mark.parametrize("k", [1, 20, 25, 32])
def test_knn_small_sets(ctx, oracle, n, k):
    rng = np.random.default_rng(n * 100 + k)
    pts = _pts(rng, n)
    ctx.nodes_reset(4, max(n, 64), W, H)
    assert ctx.nodes_append(pts) == 0
    q = np.concatenate([_pts(rng, 500).astype(np.float64), pts[: min(n, 300)].astype(np.float64)])
    _check(ctx, oracle, pts, q, k)


def test_knn_incremental_append_sees_all_points(ctx, oracle):
    rng = np.random.default_rng(3)
    ctx.nodes_reset(3, 20000, W, H)
    allp = np.zeros((0, 2), np.float32)
    for batch in (1, 7, 100, 400, 513, 3000, 1, 1, 5000):
        p = _pts(rng, batch)
        first = ctx.nodes_append(p)
        assert first == len(allp)
        allp = np.concatenate([allp, p])
        q = np.concatenate([_pts(rng, 200).astype(np.float64), p[:50].astype(np.float64)])
        _check(ctx, oracle, allp, q, 25)
    assert ctx.nodes_count() == len(allp)


def test_knn_ties_lattice(ctx, oracle):
    # integer lattice: many exactly equal squared distances -> the (d2, id) tie rule decides
    xs, ys = np.meshgrid(np.arange(0, 64, dtype=np.float32) * 4, np.arange(0, 64, dtype=np.float32) * 4)
    pts = np.stack([xs.ravel(), ys.ravel()], 1).astype(np.float32)
    rng = np.random.default_rng(5)
    pts = pts[rng.permutation(len(pts))]
    ctx.nodes_reset(4, len(pts), 256, 256)
    ctx.nodes_append(pts)
    q = np.concatenate([pts[:400].astype(np.float64), pts[:400].astype(np.float64) + 2.0])
    for k in (1, 5, 25):
        _check(ctx, oracle, pts, q, k)


def test_knn_queries_outside_domain_and_double_coords(ctx, oracle):
    rng = np.random.default_rng(9)
    pts = _pts(rng, 5000)
    ctx.nodes_reset(5, 5000, W, H)
    ctx.nodes_append(pts)
    q = np.array([[-50.0, -50.0], [W + 100.0, 10.0], [10.0, H + 500.0], [-1.0, -1.0], [W * 2.0, H * 2.0]])
    q = np.concatenate([q, rng.random((500, 2)) * np.array([W, H])])  # genuine binary64 queries (CDRT)
    _check(ctx, oracle, pts, q, 20)


def test_knn_million_nodes(ctx, oracle):
    # C4 scale: 1e6 nodes on a 4096^2 domain, 16-pixel cells
    rng = np.random.default_rng(42)
    n = 1_000_000
    pts = _pts(rng, n, 4096, 4096)
    ctx.nodes_reset(4, n, 4096, 4096)
    ctx.nodes_append(pts)
    q = np.concatenate([pts[rng.integers(0, n, 100_000)].astype(np.float64), _pts(rng, 50_000, 4096, 4096).astype(np.float64)])
    _check(ctx, oracle, pts, q, 25, grid=(4096, 4096, 16))


def test_knn_all_nodes_table(ctx, oracle):
    import ctypes as C

    rng = np.random.default_rng(77)
    n, k = 50_000, 25
    pts = _pts(rng, n, 1313, 1313)
    ctx.nodes_reset(3, n, 1313, 1313)
    ctx.nodes_append(pts)
    d_idx = ctx.dev_alloc(n * k * 4)
    d_d2 = ctx.dev_alloc(n * k * 8)
    ctx.knn_nodes_dev(k, d_idx, d_d2)
    idx = np.zeros((n, k), np.int32)
    d2 = np.zeros((n, k), np.float64)
    ctx.d2h(idx, d_idx)
    ctx.d2h(d2, d_d2)
    ctx.sync()
    ctx.dev_free(d_idx)
    ctx.dev_free(d_d2)
    oi, od, _ = oracle.knn(pts, pts.astype(np.float64), k, grid=(1313, 1313, 16), nthreads=8)
    assert np.array_equal(idx, oi)
    assert np.array_equal(d2.view(np.uint64), od.view(np.uint64))
    assert np.array_equal(idx[:, 0], np.arange(n))  # the query node itself comes back first at d2 = 0


@pytest.mark.parametrize("k", [1, 7, 25, 28, 29])
def test_knn_all_nodes_table_dense_fast_path(ctx, oracle, k):
    """The radius-filter fast path of the bulk kernels (3x3 cell block, threshold, one warp sort, certificate) and its
    fall-backs: blocks with more candidates than the fast path holds, exact duplicates and lattice points (equal
    distances -> the (d2, id) rule), nodes on the domain border, k above the fast path's limit (29)."""
    rng = np.random.default_rng(1000 + k)
    w = h = 700
    dense = _pts(rng, 60_000, w, h)                                   # ~280 candidates per 3x3 block of 16-px cells
    clump = (_pts(rng, 3000, 40, 40) + np.float32(300)).astype(np.float32)  # a block far above the candidate limit
    xs, ys = np.meshgrid(np.arange(0, 40, dtype=np.float32) * 2 + 100, np.arange(0, 40, dtype=np.float32) * 2 + 500)
    lattice = np.stack([xs.ravel(), ys.ravel()], 1).astype(np.float32)
    border = np.array([[0, 0], [0, 350.5], [699.99, 699.99], [350.25, 0], [699.5, 10]], np.float32)
    pts = np.concatenate([dense, clump, lattice, dense[:500], border])  # dense[:500] again = exact duplicates
    pts = np.ascontiguousarray(pts[rng.permutation(len(pts))])
    n = len(pts)
    ctx.nodes_reset(4, n, w, h)
    ctx.nodes_append(pts)
    d_idx = ctx.dev_alloc(n * k * 4)
    d_d2 = ctx.dev_alloc(n * k * 8)
    ctx.knn_nodes_dev(k, d_idx, d_d2)
    idx = np.zeros((n, k), np.int32)
    d2 = np.zeros((n, k), np.float64)
    ctx.d2h(idx, d_idx)
    ctx.d2h(d2, d_d2)
    ctx.sync()
    ctx.dev_free(d_idx)
    ctx.dev_free(d_d2)
    oi, od, _ = oracle.knn(pts, pts.astype(np.float64), k, grid=(w, h, 16), nthreads=8)
    assert np.array_equal(idx, oi)
    assert np.array_equal(d2.view(np.uint64), od.view(np.uint64))
    # the same through the query kernel (arbitrary binary64 queries take the uncached fast path)
    q = np.concatenate([pts[:20_000].astype(np.float64), rng.random((20_000, 2)) * np.array([w, h])])
    _check(ctx, oracle, pts, q, k, grid=(w, h, 16))

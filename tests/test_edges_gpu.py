"""K1 parity: CUDA edge checks (through the C ABI) vs the CPU oracle, bit-exact."""
import numpy as np
import pytest

from helpers import dyn_plane, edge_mix

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def game0(oracle, ctx):
    S = oracle.load_map("GAME0")
    D = dyn_plane(S.shape, np.random.default_rng(7))
    ctx.map_set_plane(1, S)
    ctx.map_set_plane(2, D)
    return S, D


def test_plane_roundtrip(ctx, game0):
    S, D = game0
    assert np.array_equal(ctx.map_get_plane(1), (S != 0).astype(np.uint8))
    assert np.array_equal(ctx.map_get_plane(2), (D != 0).astype(np.uint8))
    assert np.array_equal(ctx.map_get_plane(3), ((S != 0) | (D != 0)).astype(np.uint8))


@pytest.mark.parametrize("divisor,mask", [(5, 3), (5, 1), (2, 3), (10, 2), (2, 1)])
def test_edges_check_bit_exact(ctx, oracle, game0, divisor, mask):
    S, D = game0
    H, W = S.shape
    e = edge_mix(np.random.default_rng(100 + divisor + mask), 1_000_003, W, H)
    got = ctx.edges_check(e, divisor, mask)
    want = oracle.edges_check(S, D, e, divisor, mask, nthreads=8)
    assert got.shape == want.shape
    assert int((got != want).sum()) == 0
    assert 0.02 < want.mean() < 0.98  # the batch exercises both verdicts


def test_edges_symmetric_and_edge_cases(ctx, oracle, game0):
    S, D = game0
    H, W = S.shape
    e = edge_mix(np.random.default_rng(5), 200_000, W, H)
    fwd = ctx.edges_check(e, 5, 3)
    rev = ctx.edges_check(e[:, [2, 3, 0, 1]], 5, 3)
    assert np.array_equal(fwd, rev)  # the verdict is symmetric
    # empty, single, ragged (n not a multiple of 32), zero-length edges, far corners
    assert ctx.edges_check(np.zeros((0, 4), np.float32), 5, 3).size == 0
    special = np.array([[10, 10, 10, 10], [0, 0, W - 1, H - 1], [W - 0.5, H - 0.5, 0.25, 0.25],
                        [W - 1, 0, 0, H - 1], [550, 550, 551, 551]], np.float32)
    for n in (1, 2, 5):
        got = ctx.edges_check(special[:n], 5, 3)
        assert np.array_equal(got, oracle.edges_check(S, D, special[:n], 5, 3))
    # out-of-row / out-of-buffer coordinates follow the linear-index rule of cv::Mat::at
    oob = np.array([[W + 3.0, 10, W + 40.0, 12], [5, H + 2.0, 6, H + 9.0], [-3.5, 4, 20, 4], [W, H - 1, W, H - 1]], np.float32)
    assert np.array_equal(ctx.edges_check(oob, 5, 3), oracle.edges_check(S, D, oob, 5, 3))


def test_edges_first_hit_bit_exact(ctx, oracle, game0):
    S, D = game0
    H, W = S.shape
    e = edge_mix(np.random.default_rng(11), 300_001, W, H, short_px=150.0)
    gb, gp = ctx.edges_first_hit(e)
    wb, wp = oracle.edges_first_hit(S, D, e, nthreads=8)
    assert int((gb != wb).sum()) == 0
    assert np.array_equal(gp.view(np.uint32), wp.view(np.uint32))


def test_edges_f64_bit_exact(ctx, oracle, game0):
    S, D = game0
    H, W = S.shape
    rng = np.random.default_rng(13)
    e = edge_mix(rng, 300_000, W, H, short_px=200.0).astype(np.float64)
    e += rng.random(e.shape) * 1e-3  # genuine binary64 coordinates
    e = np.clip(e, 0, min(W, H) - 1e-9)
    got = ctx.edges_check_f64(e, 10, 2)
    want = oracle.edges_check_f64(S, D, e, 10, 2, nthreads=8)
    assert int((got != want).sum()) == 0
    assert want.any() and not want.all()


def test_dist_bit_exact(ctx, oracle):
    e = edge_mix(np.random.default_rng(17), 100_000, 4096, 4096)
    assert np.array_equal(ctx.dist(e).view(np.uint32), oracle.dist(e).view(np.uint32))


@pytest.mark.parametrize("name", ["FLOOR5", "MAZE0", "MESS6", "Trap"])
def test_edges_other_maps(ctx, oracle, name):
    S = oracle.load_map(name)
    H, W = S.shape
    ctx.map_set_plane(1, S)
    ctx.map_clear_plane(2)
    e = edge_mix(np.random.default_rng(23), 200_000, W, H)
    for divisor in (5, 2):
        got = ctx.edges_check(e, divisor, 3)
        assert np.array_equal(got, oracle.edges_check(S, None, e, divisor, 1, nthreads=8))


@pytest.mark.parametrize("n", [20_000, 45_000])
def test_tick_edges_check_equals_explicit_list(ctx, oracle, n):
    """The tick-level call over the resident tree (parents + k-NN table) returns, edge for edge, the
    verdicts of the explicit end-point list it stands for, and of the oracle.  n = 20 000 takes the gathering
    kernel, n = 45 000 (>= 2^20 edges) the materialised edge list, twice, with re-parented nodes in between."""
    S = oracle.load_map("MESS6")
    H, W = S.shape
    rng = np.random.default_rng(31)
    D = dyn_plane(S.shape, rng, n_blobs=6)
    ctx.map_set_plane(1, S)
    ctx.map_clear_plane(2)
    k = 25
    xy = (rng.random((n, 2), dtype=np.float32) * np.array([W, H], np.float32)).astype(np.float32)
    ctx.nodes_reset(4, n, W, H)
    ctx.nodes_append(xy)
    ctx.knn_table_build(k)
    idx, _, _ = ctx.knn(xy.astype(np.float64), k)
    par = rng.integers(0, n, n).astype(np.int32)
    par[0] = -1
    ctx.tree_set(par, np.zeros(n, np.float32), np.zeros(n, np.uint8))
    got = ctx.tick_edges_check(D, 2, 3)  # uploads the dynamic plane, checks (n-1) + n*k edges
    tree = np.concatenate([xy[par[1:]], xy[1:]], 1)
    rew = np.concatenate([np.repeat(xy, k, axis=0), xy[idx.reshape(-1)]], 1)
    e = np.ascontiguousarray(np.concatenate([tree, rew]), np.float32)
    assert len(got) == len(e)
    assert np.array_equal(got, ctx.edges_check(e, 2, 3))
    assert np.array_equal(got, oracle.edges_check(S, D, e, 2, 3, nthreads=8))
    # a second tick after rewires: the tree part of the list follows the parents
    par2 = rng.integers(0, n, n).astype(np.int32)
    par2[0] = -1
    par2[5] = -1  # a detached node: "no edge" -> verdict 0
    ctx.tree_set(par2, np.zeros(n, np.float32), np.zeros(n, np.uint8))
    got2 = ctx.tick_edges_check(None, 2, 3)
    e2 = e.copy()
    e2[: n - 1] = np.concatenate([xy[par2[1:]], xy[1:]], 1)
    want2 = oracle.edges_check(S, D, e2, 2, 3, nthreads=8)
    want2[4] = 0
    assert np.array_equal(got2, want2)


def _peer_alloc(c, nbytes):
    import ctypes as C

    ptr, h = C.c_void_p(), (C.c_ubyte * 64)()
    c._ck(c.L.rtpp_peer_alloc(c.h, C.c_size_t(nbytes), C.byref(ptr), h))
    return ptr.value


@pytest.mark.parametrize("world", [1, 2, 3])
def test_edges_check_allgather_protocol_on_one_gpu(oracle, world):
    """rtpp_edges_check_allgather_dev (SURVEY 8(e) grain 2) with every 'rank' on this GPU: one ctx (own stream) per rank,
    each checks its round-robin share of the blocks, stores the verdict words into every rank's buffer, defers the very
    long edges to the warp-per-edge kernel and exchanges flags; every rank's buffer must then hold exactly the verdicts
    of the oracle.  (Across processes the buffers are CUDA IPC mappings: bench.py --gpus N / tools/peer_batch_bench.py.)"""
    import ctypes as C

    import rtpp_b200

    S = oracle.load_map("MESS6")
    H, W = S.shape
    rng = np.random.default_rng(50 + world)
    D = dyn_plane(S.shape, rng)
    n = 300_001
    e = edge_mix(rng, n, W, H, long_frac=0.2)  # 60 k edges up to ~1800 px: well past the deferral threshold
    want = oracle.edges_check(S, D, e, 2, 3)
    ctxs = [rtpp_b200.Ctx(0) for _ in range(world)]
    d_edges, bits, flags = [], [], []
    try:
        nwords = (n + 31) // 32
        for c in ctxs:
            c.map_set_plane(1, S)
            c.map_set_plane(2, D)
            p = c.dev_alloc(n * 16)
            c.h2d(p, e)
            d_edges.append(p)
            bits.append([_peer_alloc(c, nwords * 4), _peer_alloc(c, nwords * 4)])
            flags.append(_peer_alloc(c, 64))
            c.sync()
        ftab = (C.c_void_p * world)(*flags)
        for epoch in (1, 2, 3):
            btab = (C.c_void_p * world)(*[b[epoch & 1] for b in bits])
            for r, c in enumerate(ctxs):  # asynchronous launches: the ranks' kernels overlap on the device
                c._ck(c.L.rtpp_edges_check_allgather_dev(c.h, C.c_void_p(d_edges[r]), C.c_size_t(n), 2, 3, btab, ftab, r, world, epoch))
            for r, c in enumerate(ctxs):
                c.sync()
            for r, c in enumerate(ctxs):
                got = np.zeros(nwords, np.uint32)
                c.d2h(got, bits[r][epoch & 1])
                c.sync()
                assert np.array_equal(rtpp_b200.Ctx.unpack_bits(got, n), want), (world, epoch, r)
    finally:
        for r, c in enumerate(ctxs):
            if r < len(flags):
                for p in (d_edges[r], bits[r][0], bits[r][1], flags[r]):
                    c.dev_free(p)
            c.close()

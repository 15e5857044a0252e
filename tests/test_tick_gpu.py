"""GPU: the windowed tick check (csrc/tick.cu: map window in shared memory by bulk async copies, supercell-ordered plan,
packed binary32x2 interpolation) against the oracle's CrashDetection on the explicit edges it stands for — bit for bit, on
bundled maps whose sizes are not multiples of the 64-pixel supercell, with nodes on the map's borders, neighbours beyond
the 32-pixel halo (general-routine fallback), fewer than k nodes, every divisor / plane mask and several hash cell sizes."""
import os

import numpy as np
import pytest

from helpers import dyn_plane

pytestmark = pytest.mark.gpu


def _knn_table(ctx, n, k):
    import ctypes as C

    p, kk = C.c_void_p(), C.c_int()
    ctx.call("knn_table_dev", C.byref(p), C.byref(kk))
    assert kk.value == k
    idx = np.zeros((n, k), np.int32)
    ctx.d2h(idx, p.value)
    ctx.sync()
    return idx


def _nodes(rng, n, W, H, S=None, border=True):
    xy = (rng.random((n, 2), dtype=np.float32) * np.array([W, H], np.float32)).astype(np.float32)
    if S is not None:  # keep them in free space like a planner's nodes (plus a few on obstacles)
        ok = S[xy[:, 1].astype(np.int32), xy[:, 0].astype(np.int32)] == 0
        ok[: n // 50] = True
        xy = xy[ok]
    if border:  # nodes on the first / last rows and columns, exact integers, the very last representable coordinate
        hiw, hih = np.nextafter(np.float32(W), np.float32(0)), np.nextafter(np.float32(H), np.float32(0))
        extra = [(0.0, 0.0), (hiw, hih), (0.0, hih), (hiw, 0.0), (W - 1.0, 10.5), (W - 1.5, 11.0), (10.0, H - 1.0), (64.0, 64.0),
                 (63.999996, 64.0), (128.0, 31.999998), (W - 2.0, H - 2.0), (W / 2, hih), (hiw, H / 2)]
        m = len(extra)
        band = rng.random((400, 2), dtype=np.float32) * np.array([W, 3.0], np.float32)
        band2 = band[:, ::-1] * np.array([1.0, H / W], np.float32)
        far = np.array([W, H], np.float32) - rng.random((400, 2), dtype=np.float32) * np.array([3.0, H], np.float32)
        xy = np.concatenate([xy, np.array(extra, np.float32), band, band2, np.clip(far, 0, [hiw, hih]).astype(np.float32)])
        assert m > 0
    return np.ascontiguousarray(xy, np.float32)


def _check_tick(ctx, oracle, S, D, xy, k, cell_log2, divisor, mask, rng, expect_usable=True):
    H, W = S.shape
    n = len(xy)
    ctx.nodes_reset(cell_log2, n, W, H)
    ctx.nodes_append(xy)
    ctx.knn_table_build(k)
    idx = _knn_table(ctx, n, k)
    par = rng.integers(0, n, n).astype(np.int32)
    par[0] = -1
    if n > 7:
        par[7] = -1
    ctx.tree_set(par, np.zeros(n, np.float32), np.zeros(n, np.uint8))
    info = ctx.tick_plan_info(divisor)
    assert info["usable"] == expect_usable, info["far_fraction"]
    if not expect_usable:
        return info
    slot_id, slot_of = info["slot_id"], info["slot_of_id"]
    assert info["n_slots"] % 32 == 0 and np.array_equal(np.sort(slot_id[slot_id >= 0]), np.arange(n))
    assert np.array_equal(slot_id[slot_of], np.arange(n))
    tree, rows = ctx.tick_check(divisor, mask)
    # tree edges: parent -> child, 0 for a detached child
    e_tree = np.concatenate([xy[np.maximum(par[1:], 0)], xy[1:]], 1)
    want_tree = oracle.edges_check(S, D, e_tree, divisor, mask, nthreads=8) if n > 1 else np.zeros(0, np.uint8)
    want_tree[par[1:] < 0] = 0
    assert np.array_equal(tree, want_tree)
    # rewire edges: bit j of the node's word
    have = idx >= 0
    e = np.concatenate([np.repeat(xy, k, axis=0), xy[np.maximum(idx, 0).reshape(-1)]], 1)
    want = oracle.edges_check(S, D, e, divisor, mask, nthreads=8).reshape(n, k).astype(bool) & have
    got = ((rows[slot_of][:, None] >> np.arange(k, dtype=np.uint32)[None, :]) & 1).astype(bool)
    bad = np.argwhere(got != want)
    if len(bad) and os.environ.get("RTPP_TICK_DUMP"):
        np.savez(os.environ["RTPP_TICK_DUMP"], bad=bad, got=got[bad[:, 0], bad[:, 1]], xy=xy, idx=idx, D=D, slot_of=slot_of, divisor=divisor, mask=mask)
    assert len(bad) == 0, (len(bad), bad[:5], xy[bad[:5, 0]], xy[idx[bad[:5, 0], bad[:5, 1]]])
    assert not rows[slot_id < 0].any()  # padding slots stay 0
    if k < 32:
        assert not (rows >> np.uint32(k)).any()
    return info


@pytest.mark.parametrize("name,n,k,cell", [("MESS6", 60_000, 25, 4), ("MESS6", 60_000, 25, 3), ("FLOOR5", 150_000, 25, 5),
                                          ("MAZE0", 200_000, 12, 4), ("MESS6", 30_000, 32, 6), ("GAME0", 120_000, 1, 4)])
def test_tick_check_bit_exact(ctx, oracle, name, n, k, cell):
    S = oracle.load_map(name)
    rng = np.random.default_rng(sum(map(ord, name)) * 1000 + n + k)
    D = dyn_plane(S.shape, rng, n_blobs=8)
    ctx.map_set_plane(1, S)
    ctx.map_set_plane(2, D)
    H, W = S.shape
    xy = _nodes(rng, n, W, H, S)
    for divisor, mask in ((2, 3), (5, 3), (5, 1), (10, 2), (1, 3)):
        info = _check_tick(ctx, oracle, S, D, xy, k, cell, divisor, mask, rng)
    assert info["far_fraction"] < 0.25


def test_tick_check_sparse_neighbours_take_the_general_routine(ctx, oracle):
    """A thinner tree: a good share of the 25-NN edges leave the 32-pixel halo and are re-checked by the general routine."""
    S = oracle.load_map("MESS6")
    rng = np.random.default_rng(77)
    D = dyn_plane(S.shape, rng, n_blobs=8)
    ctx.map_set_plane(1, S)
    ctx.map_set_plane(2, D)
    H, W = S.shape
    xy = _nodes(rng, 9_000, W, H, None)
    info = _check_tick(ctx, oracle, S, D, xy, 25, 4, 2, 3, rng)
    assert 0.01 < info["far_fraction"] < 0.25
    # thinner still: the plan declares itself unusable, the windowed entry points refuse, the legacy call still answers
    import rtpp_b200

    xy2 = _nodes(rng, 1_500, W, H, None, border=False)
    info = _check_tick(ctx, oracle, S, D, xy2, 25, 4, 2, 3, rng, expect_usable=False)
    assert info["far_fraction"] > 0.25
    with pytest.raises(rtpp_b200.RtppError):
        ctx.tick_check(2, 3)
    got = ctx.tick_edges_check(None, 2, 3)
    assert len(got) == (len(xy2) - 1) + len(xy2) * 25


def test_tick_check_fewer_nodes_than_k_and_large_cells(ctx, oracle):
    S = oracle.load_map("MESS6")
    rng = np.random.default_rng(5)
    D = dyn_plane(S.shape, rng, n_blobs=4)
    ctx.map_set_plane(1, S)
    ctx.map_set_plane(2, D)
    xy = (np.array([[700.0, 700.0]], np.float32) + rng.random((11, 2), dtype=np.float32) * 20).astype(np.float32)
    _check_tick(ctx, oracle, S, D, xy, 25, 4, 2, 3, rng)
    _check_tick(ctx, oracle, S, D, xy[:1], 25, 4, 2, 3, rng)
    # hash cells of 128 pixels are larger than a supercell: the plan does not apply
    H, W = S.shape
    big = _nodes(rng, 20_000, W, H, S)
    _check_tick(ctx, oracle, S, D, big, 25, 7, 2, 3, rng, expect_usable=False)


def test_legacy_tick_order_goes_through_the_window_kernel(ctx, oracle):
    """rtpp_tick_edges_check keeps its edge order (tree edges, then node-major rewire edges) on top of the windowed kernel."""
    S = oracle.load_map("GAME0")
    rng = np.random.default_rng(9)
    D = dyn_plane(S.shape, rng, n_blobs=8)
    H, W = S.shape
    ctx.map_set_plane(1, S)
    ctx.map_set_plane(2, D)
    xy = _nodes(rng, 100_000, W, H, S)
    n, k = len(xy), 25
    ctx.nodes_reset(4, n, W, H)
    ctx.nodes_append(xy)
    ctx.knn_table_build(k)
    idx = _knn_table(ctx, n, k)
    par = rng.integers(0, n, n).astype(np.int32)
    par[0] = -1
    ctx.tree_set(par, np.zeros(n, np.float32), np.zeros(n, np.uint8))
    assert ctx.tick_plan_info(5, want_ids=False)["usable"]
    ctx.tick_edges_check(None, 5, 3)  # the first tick also deals the groups to the launch's warps (plan_partition_kernel)
    launches = ctx.kernel_launches()
    got = ctx.tick_edges_check(None, 5, 3)
    assert ctx.kernel_launches() - launches == 2  # tick_check_kernel + tick_repack_kernel
    e = np.concatenate([np.concatenate([xy[par[1:]], xy[1:]], 1), np.concatenate([np.repeat(xy, k, axis=0), xy[idx.reshape(-1)]], 1)])
    assert np.array_equal(got, oracle.edges_check(S, D, np.ascontiguousarray(e, np.float32), 5, 3, nthreads=8))


def _peer_alloc(c, nbytes):
    import ctypes as C

    ptr, h = C.c_void_p(), (C.c_ubyte * 64)()
    c._ck(c.L.rtpp_peer_alloc(c.h, C.c_size_t(max(nbytes, 4)), C.byref(ptr), h))
    return ptr.value


@pytest.mark.parametrize("world", [2, 3])
def test_tick_allgather_protocol_on_one_gpu(ctx, oracle, world):
    """rtpp_tick_check_allgather_dev with every 'rank' on this GPU, one after the other on one stream (kernels that wait for one
    another must not share a GPU): before a rank's launch the flags of the ranks that have not run yet are set by hand, as if
    they had already signalled; the flags of the ranks that HAVE run are the ones their kernels wrote.  After the last rank every
    rank's buffers hold the whole tick, equal to the single-GPU verdicts; the work items each rank took are disjoint."""
    import ctypes as C

    S = oracle.load_map("MESS6")
    rng = np.random.default_rng(600 + world)
    D = dyn_plane(S.shape, rng, n_blobs=8)
    H, W = S.shape
    ctx.map_set_plane(1, S)
    ctx.map_set_plane(2, D)
    xy = _nodes(rng, 50_000, W, H, S)
    n, k = len(xy), 25
    ctx.nodes_reset(4, n, W, H)
    ctx.nodes_append(xy)
    ctx.knn_table_build(k)
    par = rng.integers(0, n, n).astype(np.int32)
    par[0] = -1
    ctx.tree_set(par, np.zeros(n, np.float32), np.zeros(n, np.uint8))
    want_tree, want_rows = ctx.tick_check(2, 3)
    n_slots, tree_words = len(want_rows), (n - 1 + 31) // 32
    bufs = [[_peer_alloc(ctx, tree_words * 4), _peer_alloc(ctx, n_slots * 4)] for _ in range(world)]
    flags = [_peer_alloc(ctx, 64) for _ in range(world)]
    try:
        ttab = (C.c_void_p * world)(*[b[0] for b in bufs])
        rtab = (C.c_void_p * world)(*[b[1] for b in bufs])
        ftab = (C.c_void_p * world)(*flags)
        for epoch in (1, 2):
            for r in range(world):
                later = np.zeros(16, np.int32)
                ctx.d2h(later, flags[r])
                ctx.sync()
                for q in range(r + 1, world):
                    later[q] = epoch  # rank q has "already signalled" rank r
                ctx.h2d(flags[r], later)
                ctx.call("tick_check_allgather_dev", 2, 3, ttab, rtab, ftab, r, world, epoch)
                t = C.c_int()
                ctx.call("peer_status", C.c_void_p(flags[r]), C.byref(t))
                assert t.value == 0
            for r in range(world):
                tree = np.zeros(tree_words, np.uint32)
                rows = np.zeros(n_slots, np.uint32)
                ctx.d2h(tree, bufs[r][0])
                ctx.d2h(rows, bufs[r][1])
                f = np.zeros(16, np.int32)
                ctx.d2h(f, flags[r])
                ctx.sync()
                assert np.array_equal(ctx.unpack_bits(tree, n - 1), want_tree), (world, epoch, r)
                assert np.array_equal(rows, want_rows), (world, epoch, r)
                assert all(f[q] == epoch for q in range(world) if q != r)
    finally:
        for b in bufs:
            ctx.dev_free(b[0]), ctx.dev_free(b[1])
        for f in flags:
            ctx.dev_free(f)


def test_tick_allgather_fails_closed_when_a_peer_never_arrives(ctx, oracle):
    """A rank whose peer never signals gives up after the bounded wait: rtpp_peer_status reports the timeout and the rank's
    result buffers read BLOCKED everywhere (a lost peer must never look like free edges)."""
    import ctypes as C

    S = oracle.load_map("MESS6")
    rng = np.random.default_rng(33)
    H, W = S.shape
    ctx.map_set_plane(1, S)
    ctx.map_clear_plane(2)
    xy = _nodes(rng, 30_000, W, H, S)
    n = len(xy)
    ctx.nodes_reset(4, n, W, H)
    ctx.nodes_append(xy)
    ctx.knn_table_build(25)
    par = rng.integers(0, n, n).astype(np.int32)
    par[0] = -1
    ctx.tree_set(par, np.zeros(n, np.float32), np.zeros(n, np.uint8))
    n_slots, tree_words = ctx.tick_plan_info(2, want_ids=False)["n_slots"], (n - 1 + 31) // 32
    bufs = [[_peer_alloc(ctx, tree_words * 4), _peer_alloc(ctx, n_slots * 4)] for _ in range(2)]
    flags = [_peer_alloc(ctx, 64) for _ in range(2)]
    try:
        ctx.call("peer_set_timeout_ms", C.c_double(20.0))
        ttab, rtab, ftab = ((C.c_void_p * 2)(*v) for v in ([b[0] for b in bufs], [b[1] for b in bufs], flags))
        ctx.call("tick_check_allgather_dev", 2, 3, ttab, rtab, ftab, 0, 2, 1)  # rank 1 is never launched
        t = C.c_int()
        ctx.call("peer_status", C.c_void_p(flags[0]), C.byref(t))
        assert t.value == 1
        rows = np.zeros(n_slots, np.uint32)
        tree = np.zeros(tree_words, np.uint32)
        ctx.d2h(rows, bufs[0][1])
        ctx.d2h(tree, bufs[0][0])
        ctx.sync()
        assert (rows == 0xFFFFFFFF).all() and (tree == 0xFFFFFFFF).all()
        ctx.call("peer_status", C.c_void_p(flags[0]), C.byref(t))
        assert t.value == 0  # cleared by the read
    finally:
        ctx.call("peer_set_timeout_ms", C.c_double(1000.0))
        for b in bufs:
            ctx.dev_free(b[0]), ctx.dev_free(b[1])
        for f in flags:
            ctx.dev_free(f)

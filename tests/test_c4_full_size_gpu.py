"""BASELINE configs[3] (C4) at its FULL size — 4096x4096 map, 1e6 nodes, 26.0e6 candidate edges — through the C ABI.
The oracle cannot redo 26 M edges and 1e6 25-NN queries in seconds, so the full-size run is checked through properties that
do not depend on the size: verdict symmetry and batch-order independence, oracle spot checks on strided
samples, the k-NN definition on the table's own rows, the propagate recurrence along every tree edge and the fixpoint
conditions of the bulk relax over all 25e6 rewire edges.  Same workload object as bench.py (one build, ~10 s)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def wl():
    import torch

    import bench

    prev = torch.cuda.current_stream()
    w = bench.Workload(0, seed=0)
    w.explicit_step()
    w.step()
    torch.cuda.synchronize()
    yield w
    torch.cuda.set_stream(prev)
    w.ctx.close()


def _bits(t, n):
    return np.unpackbits(t.cpu().numpy().view(np.uint8), bitorder="little")[:n]


def _dist32(a, b):
    """RTpoint::operator% (RTRRT.h:47-52) in numpy binary32, one rounding per operation."""
    dx = (a[:, 0] - b[:, 0]).astype(np.float32)
    dy = (a[:, 1] - b[:, 1]).astype(np.float32)
    return np.sqrt((dx * dx).astype(np.float32) + (dy * dy).astype(np.float32), dtype=np.float32)


def test_edge_verdicts_full_size(wl, oracle):
    import torch

    import bench

    n = wl.n_edges
    assert n == 25_999_999
    got = _bits(wl.bits, n)
    assert 0.002 < got.mean() < 0.6  # nodes lie in free space and 25-NN edges are short: ~0.8 % are blocked
    # (1) oracle on a strided sample of the whole batch (tree edges and rewire edges) and on the long tree edges
    edges = wl.edges.cpu().numpy()
    pick = np.arange(0, n, 131)
    want = oracle.edges_check(wl.S, wl.D, edges[pick], bench.DIVISOR, 3, nthreads=8)
    assert np.array_equal(got[pick], want)
    length = _dist32(edges[: bench.N_NODES - 1, 0:2], edges[: bench.N_NODES - 1, 2:4])
    longest = np.argsort(length)[-2000:]
    assert length[longest].max() > 500  # marched by the whole warp
    assert np.array_equal(got[longest], oracle.edges_check(wl.S, wl.D, edges[longest], bench.DIVISOR, 3, nthreads=8))
    # (2) the verdict is symmetric in (ps, pe) (RTRRT.cpp:214-217: t and 1 - t are both exact dyadic fractions)
    swapped = wl.edges[:, [2, 3, 0, 1]].contiguous()
    bits2 = torch.zeros_like(wl.bits)
    wl.ctx.edges_check_dev(swapped.data_ptr(), n, bench.DIVISOR, 3, bits2.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(bits2, wl.bits)
    del swapped
    # (3) a verdict does not depend on where its edge sits in the batch: reversed order, word-aligned length
    m = (n // 32) * 32
    rev = torch.flip(wl.edges[:m], dims=[0]).contiguous()
    wl.ctx.edges_check_dev(rev.data_ptr(), m, bench.DIVISOR, 3, bits2.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(_bits(bits2, m)[::-1], got[:m])
    # (4) the static plane alone never blocks more than both planes; E2-style mask 1 vs mask 3
    wl.ctx.edges_check_dev(wl.edges.data_ptr(), n, bench.DIVISOR, 1, bits2.data_ptr())
    torch.cuda.synchronize()
    assert int((bits2 & ~wl.bits).count_nonzero().item()) == 0


def test_tick_window_full_size(wl, oracle):
    """The bench's timed step — the windowed tick over the resident tree (csrc/tick.cu) — gives, for all 25 999 999 edges, the
    verdicts of the explicit end-point list (checked against the oracle above), also through the legacy edge order of
    rtpp_tick_edges_check_dev, after a change of the dynamic plane, and for divisor 5 (a second plan: other far edges)."""
    import ctypes as C

    import torch

    import bench

    n = wl.n_edges
    want = _bits(wl.bits, n)
    got = wl.tick_verdicts(wl.tree_bits.cpu().numpy(), wl.row_bits.cpu().numpy())
    assert np.array_equal(got, want)
    assert wl.plan["far_fraction"] < 0.01
    legacy = torch.zeros_like(wl.bits)
    ne = C.c_size_t()
    wl.ctx.call("tick_edges_check_dev", bench.DIVISOR, 3, C.c_void_p(legacy.data_ptr()), C.byref(ne))
    torch.cuda.synchronize()
    assert ne.value == n and torch.equal(legacy, wl.bits)
    # other obstacle poses, divisor 5: window and explicit list must still agree everywhere
    circles = [(x + 37, y - 21, r) for (x, y, r) in wl.obstacles[0]]
    wl.ctx.map_draw_dynamic(circles, wl.obstacles[1])
    bits5 = torch.zeros_like(wl.bits)
    wl.ctx.edges_check_dev(wl.edges.data_ptr(), n, 5, 3, bits5.data_ptr())
    plan5 = wl.ctx.tick_plan_info(5)
    assert plan5["usable"] and np.array_equal(plan5["slot_of_id"], wl.plan["slot_of_id"])  # the slot order does not depend on the divisor
    tree5, rows5 = torch.zeros_like(wl.tree_bits), torch.zeros_like(wl.row_bits)
    wl.ctx.tick_check_dev(5, 3, tree5.data_ptr(), rows5.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(wl.tick_verdicts(tree5.cpu().numpy(), rows5.cpu().numpy()), _bits(bits5, n))
    pick = np.arange(0, n, 257)
    D5 = wl.ctx.map_get_plane(2) * np.uint8(127)
    assert np.array_equal(_bits(bits5, n)[pick], oracle.edges_check(wl.S, D5, wl.edges.cpu().numpy()[pick], 5, 3, nthreads=8))
    wl.ctx.map_draw_dynamic(*wl.obstacles)  # restore for the tests below
    wl.step()
    torch.cuda.synchronize()


def test_knn_table_full_size(wl, oracle):
    import bench

    idx = wl.knn_idx.cpu().numpy()
    n, k = idx.shape
    assert (n, k) == (bench.N_NODES, bench.KNN)
    assert idx.min() >= 0 and idx.max() < n
    xy64 = wl.xy.astype(np.float64)
    # squared distances exactly as CGAL evaluates them (binary64, CGAL/Euclidean_distance.h:72-83)
    d = xy64[idx] - xy64[:, None, :]
    d2 = d[:, :, 0] * d[:, :, 0] + d[:, :, 1] * d[:, :, 1]
    assert np.array_equal(idx[:, 0], np.arange(n)) or (d2[:, 0] == 0).all()  # the node itself (or a duplicate point) first
    assert (np.diff(d2, axis=1) >= 0).all()                                     # ascending
    ties = np.diff(d2, axis=1) == 0
    assert (np.diff(idx, axis=1)[ties] > 0).all()                               # equal distances by id
    assert all(len(set(r)) == k for r in idx[:: n // 5000])                    # no repeated neighbour
    # the k-th distance is the true one: oracle (exact grid k-NN) on a strided sample of the rows
    pick = np.arange(0, n, n // 4000)
    oi, od, _ = oracle.knn(wl.xy, xy64[pick], k, grid=(bench.MAP_SIZE, bench.MAP_SIZE, 16))
    assert np.array_equal(idx[pick], oi)
    assert np.array_equal(d2[pick].view(np.uint64), od.view(np.uint64))


def test_propagate_and_relax_full_size(wl):
    import torch

    import bench

    ctx = wl.ctx
    n, k = bench.N_NODES, bench.KNN
    blocked = _bits(wl.bits, wl.n_edges).astype(bool)
    tree_blocked = np.concatenate([[False], blocked[: n - 1]])  # verdict of (parent(i) -> i), i >= 1
    rew_blocked = blocked[n - 1 :].reshape(n, k)
    par = wl.parent.astype(np.int64)
    cost0 = np.full(n, 3.4e38, np.float32)
    col0 = np.ones(n, np.uint8)
    cost0[0], col0[0] = 0.0, 0
    ctx.tree_set(wl.parent, cost0, col0)
    # ---- R3: cost[i] = cost[parent] + {E(parent, i), |parent - i|} along every tree edge (RTRRT.cpp:128-134)
    ctx.tree_propagate_dev(0, bench.DIVISOR, 3)
    gp, gc, gl = ctx.tree_get()
    assert np.array_equal(gp, wl.parent)
    ch = np.arange(1, n)
    w_tree = _dist32(wl.xy[par[ch]], wl.xy[ch])
    assert np.array_equal((gc[par[ch]] + w_tree).astype(np.float32).view(np.uint32), gc[ch].view(np.uint32))
    assert np.array_equal(gl[ch].astype(bool), gl[par[ch]].astype(bool) | tree_blocked[ch])
    assert gc[0] == 0 and gl[0] == 0
    # ---- R2 bulk relax (AM-RRT* rule: free edges only, AMRRT.cpp:500-504) from the all-blocked start = single-source
    # shortest paths over the free rewire edges; checked through the fixpoint conditions over all 25e6 edges
    ctx.tree_set(wl.parent, cost0, col0)
    buckets = ctx.relax_dev(wl.knn_idx.data_ptr(), k, bench.DIVISOR, 3, free_only=True)
    torch.cuda.synchronize()
    assert buckets >= 10
    rp, rc, rl = ctx.tree_get()
    idx = wl.knn_idx.cpu().numpy().astype(np.int64)
    free = rl == 0
    assert free.mean() > 0.9 and rc[0] == 0 and rp[0] == -1
    src = np.repeat(np.arange(n), k).reshape(n, k)
    w = _dist32(wl.xy[src.ravel()], wl.xy[idx.ravel()]).reshape(n, k)
    usable = (~rew_blocked) & free[:, None] & (idx != src)
    cand = (rc[:, None] + w).astype(np.float32)
    # (a) quiescence: no free edge out of a free node still improves its target
    tgt_cost = np.where(free[idx], rc[idx], np.float32(np.inf))
    assert not (usable & (cand < tgt_cost)).any()
    # (b) every reached node hangs on an edge that realises its cost, over a free rewire edge, from a free parent
    reached = np.nonzero(free)[0][1:] if free[0] else np.nonzero(free)[0]
    reached = reached[reached != 0]
    p = rp[reached].astype(np.int64)
    assert (p >= 0).all() and free[p].all()
    assert np.array_equal((rc[p] + _dist32(wl.xy[p], wl.xy[reached])).astype(np.float32).view(np.uint32), rc[reached].view(np.uint32))
    pos = (idx[p] == reached[:, None])
    assert pos.any(1).all()                                   # reached is one of its parent's k-NN
    assert not rew_blocked[p][pos].any()                      # ... over a collision-free edge
    # (c) among the parents that give the same cost the smallest id is kept (the schedule-independent tie rule):
    # incoming candidates are the edges (x -> s) of the table, grouped by target
    order = np.argsort(idx.ravel(), kind="stable")
    tgt = idx.ravel()[order]
    ok = usable.ravel()[order]
    eq = ok & (cand.ravel()[order].view(np.uint32) == rc[tgt].view(np.uint32)) & free[tgt] & (tgt != 0)
    best = np.full(n, np.iinfo(np.int64).max)
    np.minimum.at(best, tgt[eq], src.ravel()[order][eq])
    has = best != np.iinfo(np.int64).max
    assert has[reached].all()
    assert np.array_equal(best[reached], rp[reached])

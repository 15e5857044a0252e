"""K4/K5 parity: tree re-root + cost propagation (R3) and rewire relaxation (R2) vs the oracle."""
import numpy as np
import pytest

from helpers import dyn_plane

pytestmark = pytest.mark.gpu


def _free_points(rng, S, n):
    H, W = S.shape
    out = np.zeros((0, 2), np.float32)
    while len(out) < n:
        p = (rng.random((2 * n, 2), dtype=np.float32) * np.array([W, H], np.float32)).astype(np.float32)
        ok = S[p[:, 1].astype(np.int32), p[:, 0].astype(np.int32)] == 0
        out = np.concatenate([out, p[ok]])
    return np.ascontiguousarray(out[:n])


def _device_knn_table(ctx, n, k):
    d_idx = ctx.dev_alloc(n * k * 4)
    ctx.knn_nodes_dev(k, d_idx)
    idx = np.zeros((n, k), np.int32)
    ctx.d2h(idx, d_idx)
    ctx.sync()
    return d_idx, idx


def _random_tree(idx):
    """parent = nearest neighbour with a smaller id (node 0 = root); nodes without one hang off 0."""
    n, k = idx.shape
    earlier = (idx < np.arange(n)[:, None]) & (idx >= 0)
    first = np.argmax(earlier, axis=1)
    par = np.where(earlier.any(1), idx[np.arange(n), first], 0).astype(np.int32)
    par[0] = -1
    return par


@pytest.mark.parametrize("name,n,divisor", [("MESS6", 3000, 5), ("GAME0", 60000, 5), ("GAME0", 60000, 2)])
def test_propagate_and_relax(ctx, oracle, name, n, divisor):
    S = oracle.load_map(name)
    H, W = S.shape
    rng = np.random.default_rng(n + divisor)
    D = dyn_plane(S.shape, rng, n_blobs=8)
    ctx.map_set_plane(1, S)
    ctx.map_set_plane(2, D)
    xy = _free_points(rng, S, n)
    ctx.nodes_reset(4, n, W, H)
    ctx.nodes_append(xy)
    k = 25
    d_idx, idx = _device_knn_table(ctx, n, k)
    par = _random_tree(idx)
    cost0 = np.full(n, 3.4e38, np.float32)
    col0 = np.ones(n, np.uint8)
    cost0[0], col0[0] = 0.0, 0
    ctx.tree_set(par, cost0, col0)

    # ---- R3 propagate
    reached, flags = ctx.tree_propagate(0, divisor, 3, want_flags=True)
    gp, gc, gl = ctx.tree_get()
    oc, ol, oreached = oracle.tree_propagate(S, D, xy, par, 0, divisor, 3, cost0, col0)
    assert reached == oreached == n
    assert np.array_equal(gp, par)
    assert np.array_equal(gl, ol)
    assert np.array_equal(gc.view(np.uint32), oc.view(np.uint32))
    assert 0 < ol.mean() < 1
    # flags: nothing became blocked relative to the all-blocked start; every cost shrank
    assert not (flags & 1).any() and ((flags[1:] & 4) == 4).all()

    # ---- new dynamic plane: N_b capture (RT-FMT) and a second propagate
    D2 = dyn_plane(S.shape, rng, n_blobs=10)
    ctx.map_set_plane(2, D2)
    _, flags = ctx.tree_propagate(0, divisor, 3, want_flags=True)
    gp2, gc2, gl2 = ctx.tree_get()
    oc2, ol2, _ = oracle.tree_propagate(S, D2, xy, par, 0, divisor, 3, oc, ol)
    assert np.array_equal(gl2, ol2) and np.array_equal(gc2.view(np.uint32), oc2.view(np.uint32))
    assert np.array_equal((flags & 1).astype(bool), (ol2 == 1) & (ol == 0))

    # ---- R2 relax to quiescence over the k-NN graph
    sweeps = ctx.relax_dev(d_idx, k, divisor, 3)
    rp, rc, rl = ctx.tree_get()
    fc, fl, fp, osw = oracle.relax_fixpoint(S, D2, xy, idx, divisor, 3, oc2, ol2, par)
    assert sweeps >= 1 and osw >= 1
    assert np.array_equal(rl, fl)
    free = fl == 0
    assert np.array_equal(rc[free].view(np.uint32), fc[free].view(np.uint32))
    # ---- AM-RRT* rule (free edges only, AMRRT.cpp:500-504) from the propagated state
    ctx.tree_set(par, oc2, ol2)
    ctx.relax_dev(d_idx, k, divisor, 3, free_only=True)
    _, ac, al = ctx.tree_get()
    fc2, fl2, _, _ = oracle.relax_fixpoint(S, D2, xy, idx, divisor, 3, oc2, ol2, par, free_only=True)
    assert np.array_equal(al, fl2)
    assert np.array_equal(ac[fl2 == 0].view(np.uint32), fc2[fl2 == 0].view(np.uint32))
    ctx.dev_free(d_idx)


def test_reroot_matches_oracle_planner(ctx, oracle):
    """Grow a tree with the oracle's RT-RRT*, move the agent, and compare re-root + re-cost."""
    S = oracle.load_map("MESS6")
    H, W = S.shape
    orc = oracle.OrcRTRRT(S)
    orc.set_clock(2000)
    orc.set_agent(720, 720)
    r = oracle.rand31(5, 4000000)
    pos = 0
    for _ in range(6):
        orc.update_tree()
        pts, offs = orc.presample(r[pos:], 30000)
        used = orc.expand_rewiring(50, pts)
        pos += int(offs[used])
    xy, par, cost, col = orc.dump()
    n = len(xy)
    assert n > 300
    ctx.map_set_plane(1, S)
    ctx.map_clear_plane(2)
    ctx.nodes_reset(3, n, W, H)
    ctx.nodes_append(xy)
    ctx.tree_set(par, cost, col)
    # move the agent next to some far node; UpdateTree re-roots at the nearest node
    target = xy[n // 2]
    orc.set_agent(float(target[0]), float(target[1]))
    D = dyn_plane(S.shape, np.random.default_rng(2), n_blobs=5)
    orc.set_dobs(D)
    ctx.map_set_plane(2, D)
    orc.update_tree()
    idx, _, _ = ctx.knn(np.array([target], np.float64), 1)
    new_root = int(idx[0, 0])
    old_root = int(np.flatnonzero(par < 0)[0])
    ctx.tree_reroot(old_root, new_root)
    ctx.tree_propagate(new_root, 5, 3)
    gp, gc, gl = ctx.tree_get()
    oxy, opar, ocost, ocol = orc.dump()
    assert np.array_equal(oxy, xy)
    assert np.array_equal(gp, opar)
    assert np.array_equal(gl, ocol)
    assert np.array_equal(gc.view(np.uint32), ocost.view(np.uint32))


def test_propagate_fan_out_shapes(ctx, oracle):
    """R3 on a tree whose nodes have 0, a few, hundreds and thousands of children (thread-, warp- and grid-expanded
    frontier nodes of tree_bfs_kernel) plus a long chain: costs, flags and the reached count vs the oracle."""
    S = oracle.load_map("MESS6")
    H, W = S.shape
    rng = np.random.default_rng(77)
    D = dyn_plane(S.shape, rng, n_blobs=6)
    ctx.map_set_plane(1, S)
    ctx.map_set_plane(2, D)
    n = 20000
    xy = _free_points(rng, S, n)
    par = np.zeros(n, np.int32)
    par[0] = -1
    par[1:6001] = 0                                  # 6000 children of the root (grid-expanded)
    par[6001:7501] = 1                               # 1500 children of node 1 (warp-expanded)
    par[7501:7521] = 2                               # 20 children of node 2 (warp-expanded)
    par[7521:7529] = 3                               # 8 children of node 3 (thread-expanded)
    par[7529:9529] = np.arange(7528, 9528)           # a 2000-node chain
    par[9529:] = rng.integers(1, 9529, n - 9529)     # the rest hang anywhere below
    par[15000:15010] = -1                            # detached nodes are never reached
    cost0 = np.full(n, 3.4e38, np.float32)
    col0 = np.ones(n, np.uint8)
    cost0[0], col0[0] = 0.0, 0
    ctx.nodes_reset(4, n, W, H)
    ctx.nodes_append(xy)
    ctx.tree_set(par, cost0, col0)
    reached, flags = ctx.tree_propagate(0, 5, 3, want_flags=True)
    gp, gc, gl = ctx.tree_get()
    oc, ol, oreached = oracle.tree_propagate(S, D, xy, par, 0, 5, 3, cost0, col0)
    assert reached == oreached
    assert np.array_equal(gp, par)
    assert np.array_equal(gl, ol)
    assert np.array_equal(gc.view(np.uint32), oc.view(np.uint32))


def test_relax_lattice_ties_schedule_independent(ctx, oracle):
    """Integer lattice points: nearly every node has several parents of exactly equal cost.  The work-list schedule
    (cost buckets, block-local chasing, asynchronous re-queueing) must still reach the oracle's costs and flags, and
    the tie rule (smaller parent id) makes the parents reproducible from run to run."""
    S = oracle.load_map("MESS6")
    H, W = S.shape
    ctx.map_set_plane(1, S)
    ctx.map_clear_plane(2)
    gx, gy = np.meshgrid(np.arange(8, W - 8, 6), np.arange(8, H - 8, 6))
    pts = np.stack([gx.ravel(), gy.ravel()], 1).astype(np.float32)
    pts = pts[S[pts[:, 1].astype(np.int32), pts[:, 0].astype(np.int32)] == 0]
    rng = np.random.default_rng(5)
    xy = np.ascontiguousarray(pts[rng.permutation(len(pts))])
    n, k = len(xy), 9
    assert n > 5000
    ctx.nodes_reset(4, n, W, H)
    ctx.nodes_append(xy)
    d_idx, idx = _device_knn_table(ctx, n, k)
    par = _random_tree(idx)
    cost0 = np.full(n, 3.4e38, np.float32)
    col0 = np.ones(n, np.uint8)
    cost0[0], col0[0] = 0.0, 0
    runs = []
    for free_only in (True, False):
        for _ in range(2):
            ctx.tree_set(par, cost0, col0)
            buckets = ctx.relax_dev(d_idx, k, 5, 1, free_only=free_only)
            assert buckets >= 1
            runs.append(ctx.tree_get())
        (p1, c1, l1), (p2, c2, l2) = runs[-2:]
        assert np.array_equal(p1, p2) and np.array_equal(c1.view(np.uint32), c2.view(np.uint32)) and np.array_equal(l1, l2)
        fc, fl, _, _ = oracle.relax_fixpoint(S, None, xy, idx, 5, 1, cost0, col0, par, free_only=free_only)
        assert np.array_equal(l1, fl)
        free = fl == 0
        assert free.sum() > n // 2
        assert np.array_equal(c1[free].view(np.uint32), fc[free].view(np.uint32))
        # every free node's parent realises its cost: cost[parent] + |parent - node| == cost[node] (binary32)
        ch = np.nonzero(free & (np.arange(n) != 0))[0]
        d = np.sqrt(((xy[ch] - xy[p1[ch]]) ** 2).sum(1, dtype=np.float32), dtype=np.float32)
        assert np.array_equal((c1[p1[ch]] + d).astype(np.float32).view(np.uint32), c1[ch].view(np.uint32))
    ctx.dev_free(d_idx)


@pytest.mark.parametrize("free_only", [False, True])
def test_relax_ragged_tables_and_tiny_graphs(ctx, oracle, free_only):
    """Ragged inputs of the bulk relax: a single node, rows with missing neighbours (-1), self loops, nodes nobody
    points at, a graph cut in two by a wall (the far side keeps its colliding start cost under the AM-RRT* rule and
    gets a colliding cost through the wall under the RT-RRT* rule), and costs that are already a fixpoint."""
    H = W = 256
    S = np.zeros((H, W), np.uint8)
    S[:, 120:136] = 255  # a wall from top to bottom
    ctx.map_set_plane(1, S)
    ctx.map_clear_plane(2)
    rng = np.random.default_rng(11)

    def run(xy, idx, par, cost0, col0):
        n, k = idx.shape
        ctx.nodes_reset(4, max(n, 1), W, H)
        ctx.nodes_append(xy)
        d_idx = ctx.dev_alloc(max(idx.nbytes, 4))
        ctx.h2d(d_idx, np.ascontiguousarray(idx))
        ctx.tree_set(par, cost0, col0)
        ctx.sync()
        buckets = ctx.relax_dev(d_idx, k, 5, 1, free_only=free_only)
        got = ctx.tree_get()
        fc, fl, _, _ = oracle.relax_fixpoint(S, None, xy, idx, 5, 1, cost0, col0, par, free_only=free_only)
        assert np.array_equal(got[2], fl)
        free = fl == 0
        assert np.array_equal(got[1][free].view(np.uint32), fc[free].view(np.uint32))
        # a second call from the fixpoint changes nothing
        ctx.relax_dev(d_idx, k, 5, 1, free_only=free_only)
        again = ctx.tree_get()
        assert all(np.array_equal(a, b) for a, b in zip(got, again))
        ctx.dev_free(d_idx)
        return got, buckets

    # one node, no neighbours
    got, _ = run(np.array([[10.5, 10.5]], np.float32), np.full((1, 3), -1, np.int32), np.array([-1], np.int32),
                 np.zeros(1, np.float32), np.zeros(1, np.uint8))
    assert got[0][0] == -1 and got[1][0] == 0 and got[2][0] == 0
    # two sides of the wall, ragged rows
    n, k = 400, 6
    xy = np.concatenate([rng.random((n // 2, 2)) * [100, 250] + [5, 3], rng.random((n // 2, 2)) * [100, 250] + [145, 3]]).astype(np.float32)
    d = ((xy[:, None, :].astype(np.float64) - xy[None, :, :]) ** 2).sum(2)
    idx = np.argsort(d, axis=1, kind="stable")[:, :k].astype(np.int32)   # column 0 = the node itself (self loop)
    idx[rng.random((n, k)) < 0.25] = -1                                     # missing entries anywhere in a row
    idx[7] = -1                                                             # a node without any out-edge
    idx[idx == 9] = -1                                                      # a node nobody points at
    # bridges through the wall (blocked edges)
    idx[3, 1], idx[5, 2] = n // 2 + 4, n // 2 + 8
    par = np.full(n, 0, np.int32)
    par[0] = -1
    cost0 = np.full(n, 3.4e38, np.float32)
    col0 = np.ones(n, np.uint8)
    cost0[0], col0[0] = 0.0, 0
    got, buckets = run(xy, idx, par, cost0, col0)
    assert buckets >= 1
    left_free = got[2][: n // 2] == 0
    assert left_free.sum() > n // 4
    assert (got[2][n // 2 :] == 1).all()  # nothing on the far side is reachable without crossing the wall
    if free_only:
        assert np.array_equal(got[1][n // 2 :].view(np.uint32), cost0[n // 2 :].view(np.uint32))


def test_relax_coincident_nodes_keep_the_tree_acyclic(ctx, oracle):
    """Duplicated and nearly coincident nodes: the edge between them has a weight the binary32 sum absorbs
    (cost + w == cost).  Such an edge may carry a strict cost improvement (as in the reference's sequential loop,
    RTRRT.cpp:311) but not the equal-cost parent-id tie-break, or two coincident nodes would adopt each other as
    parents: costs equal the sequential fixpoint and every free node still reaches the root by its parent links."""
    H = W = 512
    S = np.zeros((H, W), np.uint8)
    ctx.map_set_plane(1, S)
    ctx.map_clear_plane(2)
    rng = np.random.default_rng(3)
    base = (rng.random((600, 2)) * 60 + 66).astype(np.float32)     # a cloud ~ 600 px from the root: ulp(cost) = 6.1e-5
    near = base[:150].copy()
    near[:, 0] = np.nextafter(near[:, 0], np.float32(1e9))          # one coordinate ulp (7.6e-6) away: w < ulp(cost) / 2
    xy = np.ascontiguousarray(np.concatenate([[[500.5, 500.5]], base, base[300:450], near]).astype(np.float32))
    n, k = len(xy), 16
    ctx.nodes_reset(4, n, W, H)
    ctx.nodes_append(xy)
    d_idx, idx = _device_knn_table(ctx, n, k)
    # the worst start for the tie rule: every node hangs off the LAST node, so any smaller id looks better at equal cost
    par = np.full(n, n - 1, np.int32)
    par[0] = -1
    cost0 = np.full(n, 3.4e38, np.float32)
    col0 = np.ones(n, np.uint8)
    cost0[0], col0[0] = 0.0, 0
    for free_only in (True, False):
        ctx.tree_set(par, cost0, col0)
        ctx.relax_dev(d_idx, k, 5, 1, free_only=free_only)
        p, c, l = ctx.tree_get()
        fc, fl, _, _ = oracle.relax_fixpoint(S, None, xy, idx, 5, 1, cost0, col0, par, free_only=free_only)
        assert np.array_equal(l, fl)
        free = fl == 0
        assert free.sum() > n // 2
        assert np.array_equal(c[free].view(np.uint32), fc[free].view(np.uint32))
        cur = np.nonzero(free)[0]
        for _ in range(n):
            cur = cur[cur != 0]
            if len(cur) == 0:
                break
            assert (c[p[cur]] <= c[cur]).all()
            cur = np.unique(p[cur])
        assert len(cur) == 0                    # everybody arrived at the root within n hops: no cycle
    ctx.dev_free(d_idx)

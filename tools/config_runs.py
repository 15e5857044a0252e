"""BASELINE.json's configured workloads at their stated sizes, shared by tests/test_configs_gpu.py and bench.py:
  C1  RT-RRT* on bundled GAME0, 20 000 injected samples, then the 2^24 (node, one of its 25-NN) edge batch and the 2^20
      long-edge batch (E1: divisor 5, both planes);
  C3  RT-FMT on GAME0 with 1e5 free-space samples until V_unv is exhausted (or no z is left).
(C2 = tools/cdrt_bench.py at 600 ticks, C4 = bench.py's step, C5 = bench.measure_sweep.)  Every run returns the GPU numbers AND
the parity flag against the CPU oracle port of the same loop on the same injected streams."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def c1_rtrrt(ctx, orc, n_samples=20_000, dt=1_000_000, max_ticks=40_000, seed=1, check_cpu=True, check_ticks=None):
    """RT-RRT* single agent on GAME0 (RTRRT/test/test.cpp:60-67: agent (550, 550), no goal => uniform sampling only,
    RTRRT.cpp:234-239), samples injected from rand31(seed) exactly as the uniform branch draws them; ticks of
    UpdateTree + ExpandRewiring(50 ms virtual) + GetPath until the samples are consumed.  The virtual clock advances dt = 1 ms per
    read, i.e. 50 loop iterations (one sample, then queue pops) per tick: once the tree is larger than the budget a tick takes ONE
    sample whatever the budget is (the rewire-from-root sweep, RTRRT.cpp:323-353, eats the rest), so the budget only sets how
    much rewiring rides along; 20 000 samples are then 20 000 ticks.  check_ticks: compare the trees after that many ticks (a
    bounded CPU run, for bench.py) instead of at the end (tests/test_configs_gpu.py)."""
    import rtpp_b200
    import sweep_bench as sb

    S = orc.load_map("GAME0")
    H, W = S.shape
    samples = sb.uniform_samples(orc, seed, n_samples, W, H)
    eng = rtpp_b200.Engine(ctx, 1, 16384, 5, W, H, n_samples)
    eng.set_map(0, 1, S)
    eng.init_instances(np.zeros(1, np.int32), np.array([[550.0, 550.0]], np.float32), np.array([[-1.0, -1.0]], np.float32))
    eng.set_inputs(samples=samples[None], n_samples=np.array([n_samples], np.int32))
    ctx.sync()
    lat, snap = [], None
    while len(lat) < max_ticks:
        t0 = time.perf_counter()
        eng.tick(update_tree=True, run_ms=50, dt_ns=dt, get_path=True)
        ctx.sync()
        lat.append(time.perf_counter() - t0)
        if check_ticks is not None and len(lat) == check_ticks:
            snap = (eng.read_tree(0), int(eng.stats()[0, 2]))
        if int(eng.stats()[0, 2]) >= n_samples:
            break
    st = eng.stats_dict(0)
    tree = eng.read_tree(0)
    eng.close()
    if check_ticks is None or len(lat) < check_ticks:
        snap, check_ticks = (tree, int(st["samples_used"])), len(lat)
    out = {"map": "GAME0", "samples": n_samples, "ticks": len(lat), "virtual_dt_ns": dt, "nodes": int(st["n_nodes"]), "samples_used": int(st["samples_used"]),
           "gpu_status": int(st["status"]), "gpu_tick_ms_p50": float(np.median(lat) * 1e3), "gpu_total_s": float(sum(lat)),
           "gpu_iters_per_s": st["iters"] / sum(lat), "gpu_queue_pops_per_s": (st["pops_r"] + st["pops_s"]) / sum(lat)}
    if check_cpu:
        o = orc.OrcRTRRT(S)
        o.set_clock(dt)
        o.set_agent(550.0, 550.0)
        o.set_goal(-1.0, -1.0)
        pos, tl = 0, []
        for _ in range(check_ticks):
            t0 = time.perf_counter()
            o.update_tree()
            pos += o.expand_rewiring(50, samples[pos:])
            o.get_path()
            tl.append(time.perf_counter() - t0)
        cpu_tree = o.dump()
        o.close()
        out["cpu_tick_ms_p50"] = float(np.median(tl) * 1e3)
        out["cpu_total_s"] = float(sum(tl))
        out["cpu_kind"] = "port (oracle loops, grid k-NN), 1 thread"
        # the same ticks with the reference's index behaviour: a kd-tree rebuilt on the first query after every insert (CGAL,
        # Kd_tree.h:433-466) — same tree, the reference-shaped cost (report both, state which)
        o = orc.OrcRTRRT(S)
        o.knn_mode(1)
        o.set_clock(dt)
        o.set_agent(550.0, 550.0)
        o.set_goal(-1.0, -1.0)
        pos2, tk = 0, []
        for _ in range(check_ticks):
            t0 = time.perf_counter()
            o.update_tree()
            pos2 += o.expand_rewiring(50, samples[pos2:])
            o.get_path()
            tk.append(time.perf_counter() - t0)
        kd_tree = o.dump()
        out["cpu_kdtree_rebuild_tick_ms_p50"] = float(np.median(tk) * 1e3)
        out["cpu_kdtree_rebuild_total_s"] = float(sum(tk))
        out["cpu_kdtree_rebuilds"] = o.knn_mode(-1)
        out["cpu_kdtree_same_tree"] = bool(pos2 == pos and all(np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))
                                                                for a, b in zip(cpu_tree, kd_tree)))
        o.close()
        out["ticks_compared"] = check_ticks
        out["tree_matches_cpu"] = bool(pos == snap[1] and all(np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))
                                                              for a, b in zip(cpu_tree, snap[0])))
    return out, tree, S


def c1_edge_batches(ctx, orc, tree_xy, S, n_knn_edges=1 << 24, n_long=1 << 20, seed=2, check="all", timed=None):
    """The C1 micro-batches: 2^24 edges drawn as (node, one of its 25 nearest) from the 20 000-sample tree and
    2^20 long edges with uniform end points, E1 (divisor 5, S|D) through rtpp_edges_check_dev; verdicts against the oracle
    (check = "all" or a stride); timed(fn) -> ms, when given, times the resident kernel for the roofline."""
    import torch

    H, W = S.shape
    rng = np.random.default_rng(seed)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import dyn_plane

    D = dyn_plane(S.shape, rng, n_blobs=8)
    ctx.map_set_plane(1, S)
    ctx.map_set_plane(2, D)
    n = len(tree_xy)
    idx, _, _ = orc.knn(tree_xy, tree_xy.astype(np.float64), 25, grid=(W, H, 16), nthreads=os.cpu_count() or 1)
    src = rng.integers(0, n, n_knn_edges)
    dst = idx[src, rng.integers(0, 25, n_knn_edges)]
    knn_edges = np.ascontiguousarray(np.concatenate([tree_xy[src], tree_xy[np.maximum(dst, 0)]], 1), np.float32)
    long_edges = np.ascontiguousarray((rng.random((n_long, 4), dtype=np.float32) * np.array([W, H, W, H], np.float32)).astype(np.float32))
    out = {}
    dev = torch.device("cuda", torch.cuda.current_device())
    for name, e in (("knn_edges_2p24", knn_edges), ("long_edges_2p20", long_edges)):
        d_e = torch.from_numpy(e).to(dev)
        d_b = torch.zeros((len(e) + 31) // 32, dtype=torch.int32, device=dev)
        ctx.edges_check_dev(d_e.data_ptr(), len(e), 5, 3, d_b.data_ptr())
        ctx.sync()
        got = np.unpackbits(d_b.cpu().numpy().view(np.uint8), bitorder="little")[: len(e)]
        pick = slice(None) if check == "all" else slice(0, None, int(check))
        want = orc.edges_check(S, D, e[pick], 5, 3, nthreads=os.cpu_count() or 1)
        probes = orc.edges_nprobes_sum(e[:: max(1, len(e) // 1_000_000)], 5) / len(e[:: max(1, len(e) // 1_000_000)])
        r = {"edges": len(e), "blocked_fraction": float(got.mean()), "mean_probes_per_edge": float(probes), "verdicts_match_cpu": bool(np.array_equal(got[pick], want)),
             "verdicts_compared": int(len(want))}
        if timed is not None:
            ms = timed(lambda: ctx.edges_check_dev(d_e.data_ptr(), len(e), 5, 3, d_b.data_ptr()))
            r.update({"ms": ms, "edges_per_s": len(e) / (ms * 1e-3), "algorithmic_GBs": len(e) * (16 + 4 * probes + 0.125) / (ms * 1e-3) / 1e9})
        out[name] = r
        del d_e, d_b
    return out


def c3_rtfmt(ctx, orc, samples=100_000, chunk=20_000, max_rounds=2_000_000, check_cpu=True, check_rounds=None):
    """RT-FMT on GAME0 with 1e5 free-space samples (RTFMT/test/test.cpp:56 agent (360, 540); the reference ships 2000): rounds
    of ExpandFMT + RewireFromObstacles + RewireFromRoot (RTFMT.cpp:423-703) until V_unv is exhausted or no z is left."""
    import rtpp_b200

    S = orc.load_map("GAME0")
    H, W = S.shape
    o = orc.OrcRTFMT(S)
    o.set_agent(360.0, 540.0)
    o.set_goal(2000.0, 400.0)
    t0 = time.perf_counter()
    o.start(orc.rand31(1, 3 * samples + 1000), samples)
    cpu_start = time.perf_counter() - t0
    xy = o.dump()[0]
    eng = rtpp_b200.FmtEngine(ctx, 1, len(xy) + 32, 5, W, H)
    eng.set_map(0, 1, S)
    t0 = time.perf_counter()
    eng.start(np.zeros(1, np.int32), [xy], np.array([[2000.0, 400.0]], np.float32))
    ctx.sync()
    gpu_start = time.perf_counter() - t0
    rounds, gpu_s, first, snap = 0, 0.0, True, None
    while rounds < max_rounds:
        t0 = time.perf_counter()
        eng.tick(update_tree=first, iters=chunk, with_rewire=True, get_path=False)
        ctx.sync()
        gpu_s += time.perf_counter() - t0
        rounds += chunk
        first = False
        if check_rounds is not None and rounds == check_rounds:
            snap = eng.read_tree(0)
        st = eng.stats_dict(0)
        if st["n_unv"] == 0 or st["z"] < 0:
            break
    st = eng.stats_dict(0)
    out = {"samples": samples, "nodes": int(st["n_nodes"]), "rounds": rounds, "n_unv_left": int(st["n_unv"]), "z": int(st["z"]), "gpu_status": int(st["status"]),
           "gpu_start_ms": gpu_start * 1e3, "gpu_s": gpu_s, "gpu_rounds_per_s": rounds / gpu_s, "cpu_start_ms": cpu_start * 1e3}
    if check_cpu:
        if snap is None:
            snap, check_rounds = eng.read_tree(0), rounds
        t0 = time.perf_counter()
        o.update_tree()
        for _ in range(check_rounds // chunk):
            o.expand(chunk, True)
        cpu_s = time.perf_counter() - t0
        par, c, cl, stt = snap
        oxy, opar, oc, ocl, ost = o.dump()
        out["cpu_s"] = cpu_s
        out["rounds_compared"] = check_rounds
        out["cpu_rounds_per_s"] = check_rounds / cpu_s
        out["cpu_kind"] = "port (oracle loops), 1 thread"
        out["state_matches_cpu"] = bool(np.array_equal(par, opar) and np.array_equal(stt, ost) and np.array_equal(cl, ocl) and
                                        np.array_equal(c.view(np.uint32), oc.view(np.uint32)))
    eng.close()
    o.close()
    return out

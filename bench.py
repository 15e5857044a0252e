#!/usr/bin/env python
"""bench.py — throughput of the planner hot path on B200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[3], the largest single-GPU configuration): AM-RRT* on a synthetic
4096x4096 occupancy map (PCG64 seed 42: 400 rectangles + 200 discs, plus 6 dynamic obstacles) with
1e6 tree nodes.  One STEP = one batched collision pass over every candidate edge of one rewire
tick: the 999 999 tree edges (R3 re-cost, AMRRT.cpp:185-199) + the 25e6 rewire edges node ->
25-NN (AMRRT.cpp:496-500), divisor 2, both planes: 25 999 999 edges.
metric = edge collision checks / s.
`value`: the tick over the RESIDENT tree (rtpp_tick_check_dev: windowed kernel, csrc/tick.cu), K launches timed with CUDA
events on the launching stream; the launch reads ~250 MB of plan + tree (> L2) and no edge list.
N > 1 (torchrun): ONE tick — the same 25 999 999 edges — split over the ranks (strong scaling): the tree is replicated,
the tick's work items go round-robin over the ranks and every rank ends with ALL verdict words
(rtpp_tick_check_allgather_dev: the all-gather is peer stores over NVLink inside the launch); time = max over ranks.
`e2e`: one replanning tick through the host-buffer C-ABI calls a planner class makes
(rtpp_map_draw_dynamic + rtpp_tick_check): the obstacle poses go in, the verdict words come back to page-locked host
memory, the tree stays resident.  With N > 1 every rank uploads the poses and checks its share; rank 0 — the planner process —
reads the whole tick's verdict words back (the variant in which every rank reads everything back is reported beside it).  The
same edges as an explicit end-point list from the host (rtpp_edges_check, 416 MB H2D per step) are reported too.
`roofline`: algorithmic bytes of the step (16 + 4 * probes + 1/8 per edge) / launch time against the measured
HBM peak, with the physically moved bytes (ncu) and the instruction-issue form next to it.
Extras (the explicit-list kernel, the other edge kernels, k-NN queries/s, propagate / relax latency, the four planners'
replan iterations/s and tick latencies against the CPU port) are measured outside the timed region.

--impl reference: the reference's own AMRRTplanner::CrashDetection (compiled in place into oracle/_ref; the oracle port if
that library is absent) over the SAME 25 999 999 edges per step on all host threads (rank 0 only).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MAP_SIZE = 4096
N_NODES = 1_000_000
KNN = 25
DIVISOR = 2
METRIC = "edge_collision_checks_per_s"
UNIT = "edges/s"
SCALING = "strong"  # N > 1: ONE step (the same 26.0 M edges) split over the ranks


# ------------------------------------------------------------------------------------ workload
def c4_map(seed=42):
    """Synthetic C4 planes: static (seed 42) + a few dynamic discs/boxes (value 127)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    S = np.zeros((MAP_SIZE, MAP_SIZE), np.uint8)
    for _ in range(400):
        w, h = rng.integers(16, 257, 2)
        x, y = rng.integers(0, MAP_SIZE - 16, 2)
        S[y:y + h, x:x + w] = 255
    yy, xx = np.ogrid[:MAP_SIZE, :MAP_SIZE]

    def disc(img, x, y, r, v):
        y0, y1, x0, x1 = max(0, y - r), min(MAP_SIZE, y + r + 1), max(0, x - r), min(MAP_SIZE, x + r + 1)
        sub = (yy[y0:y1] - y) ** 2 + (xx[:, x0:x1] - x) ** 2 <= r * r
        img[y0:y1, x0:x1][sub] = v

    for _ in range(200):
        r = int(rng.integers(8, 97))
        x, y = (int(v) for v in rng.integers(0, MAP_SIZE, 2))
        disc(S, x, y, r, 255)
    # dynamic obstacles as the planners hold them (DObstacles): 4 discs + 2 polygons, integer pixel points
    circles = [(int(rng.integers(200, MAP_SIZE - 200)), int(rng.integers(200, MAP_SIZE - 200)), int(rng.integers(40, 120))) for _ in range(4)]
    polys = []
    for _ in range(2):
        cx, cy = (int(v) for v in rng.integers(300, MAP_SIZE - 300, 2))
        polys.append(np.array([(cx - 90, cy - 60), (cx + 90, cy - 60), (cx + 120, cy + 70), (cx, cy + 110), (cx - 120, cy + 70)], np.int32))
    return S, (circles, polys)


def c4_nodes(S, seed, n=N_NODES):
    """n uniform free-space points, float32 like rand()/float(RAND_MAX)*shape."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = []
    have = 0
    while have < n:
        p = (rng.random((n, 2), dtype=np.float32) * np.float32(MAP_SIZE)).astype(np.float32)
        ok = S[p[:, 1].astype(np.int32), p[:, 0].astype(np.int32)] == 0
        out.append(p[ok])
        have += int(ok.sum())
    return np.ascontiguousarray(np.concatenate(out)[:n])


class Workload:
    """Resident C4 state on one GPU + the candidate-edge batch."""

    def __init__(self, device, seed):
        import torch

        import rtpp_b200

        self.torch = torch
        self.dev = torch.device("cuda", device)
        torch.cuda.set_device(self.dev)
        self.stream = torch.cuda.Stream(self.dev)
        torch.cuda.set_stream(self.stream)
        self.ctx = rtpp_b200.Ctx(device)
        self.ctx.set_stream(self.stream.cuda_stream)
        self.S, self.obstacles = c4_map(42)
        self.ctx.map_set_plane(1, self.S)
        self.ctx.map_draw_dynamic(*self.obstacles)          # D1 on the device
        self.D = self.ctx.map_get_plane(2) * np.uint8(127)   # the byte plane the CPU arm probes
        self.xy = c4_nodes(self.S, 1000 + seed)
        n = len(self.xy)
        self.ctx.nodes_reset(4, n, MAP_SIZE, MAP_SIZE)
        self.ctx.nodes_append(self.xy)
        self.ctx.nodes_build()
        self.knn_idx = torch.empty((n, KNN), dtype=torch.int32, device=self.dev)
        self.ctx.knn_nodes_dev(KNN, self.knn_idx.data_ptr())
        self.ctx.sync()
        xy_t = torch.from_numpy(self.xy).to(self.dev)
        idx = self.knn_idx.long()
        # random-tree order: parent = nearest neighbour with a smaller id, node 0 = root
        ar = torch.arange(n, device=self.dev)[:, None]
        earlier = (idx < ar) & (idx >= 0)
        first = torch.argmax(earlier.int(), dim=1)
        par = torch.where(earlier.any(1), idx[ar[:, 0], first], torch.zeros_like(first))
        par[0] = -1
        self.parent = par.int().cpu().numpy()
        tree = torch.cat([xy_t[par[1:]], xy_t[1:]], 1)                       # (parent, child)
        rew = torch.cat([xy_t[:, None, :].expand(n, KNN, 2), xy_t[idx]], 2).reshape(-1, 4)  # (x, knn[x][j])
        self.edges = torch.cat([tree, rew], 0).contiguous()
        self.n_edges = int(self.edges.shape[0])
        self.bits = torch.zeros((self.n_edges + 31) // 32, dtype=torch.int32, device=self.dev)
        del xy_t, idx, earlier, rew, tree
        # the same edges as a resident tree: parents + in-ctx k-NN table (used by the e2e tick call)
        cost0 = np.full(n, 3.4e38, np.float32)
        col0 = np.ones(n, np.uint8)
        cost0[0], col0[0] = 0.0, 0
        self.ctx.tree_set(self.parent, cost0, col0)
        self.ctx.knn_table_build(KNN)
        # the tick's result buffers in the windowed kernel's layout: packed tree bits + one verdict word per plan slot
        self.plan = self.ctx.tick_plan_info(DIVISOR)
        if not self.plan["usable"]:
            raise SystemExit("the windowed tick plan does not apply to the C4 workload")
        self.tree_bits = torch.zeros((n - 1 + 31) // 32, dtype=torch.int32, device=self.dev)
        self.row_bits = torch.zeros(self.plan["n_slots"], dtype=torch.int32, device=self.dev)
        torch.cuda.synchronize()

    def step(self):
        """The timed step: one tick's collision pass over the resident tree (windowed kernel)."""
        self.ctx.tick_check_dev(DIVISOR, 3, self.tree_bits.data_ptr(), self.row_bits.data_ptr())

    def explicit_step(self):
        """The same 25 999 999 edges as an explicit end-point list (streaming kernel K1)."""
        self.ctx.edges_check_dev(self.edges.data_ptr(), self.n_edges, DIVISOR, 3, self.bits.data_ptr())

    def tick_verdicts(self, tree_words, row_words):
        """(tree words, slot row words) as numpy uint32 -> one verdict per edge in the explicit list's order."""
        n = len(self.xy)
        tb = np.unpackbits(np.ascontiguousarray(tree_words).view(np.uint8), bitorder="little")[: n - 1]
        r = np.ascontiguousarray(row_words).view(np.uint32)[self.plan["slot_of_id"]]
        rb = ((r[:, None] >> np.arange(KNN, dtype=np.uint32)[None, :]) & 1).astype(np.uint8).reshape(-1)
        return np.concatenate([tb, rb])


# ------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Polls SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.stop_flag = [], False
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        nv = self.nv
        while self.ok and not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), sm, int(reasons)))
            except Exception:
                pass
            time.sleep(self.period)

    def summary(self, t0, t1):
        nv = self.nv
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml_unavailable"]}
        inside = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples[-3:]
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                 0x80: "hw_power_brake", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}
        mask = 0
        for s in inside:
            mask |= s[2]
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception:
            mx = None
        return {"sm_mhz": float(np.median([s[1] for s in inside])), "sm_max_mhz": mx,
                "reasons": sorted(v for k, v in names.items() if mask & k), "samples": len(inside)}


# ------------------------------------------------------------------------------------ CPU side
def cpu_edge_fn(S, D, threads):
    """The reference CPU implementation of the step's edge routine on `threads` host threads -> (fn(edges) -> verdicts, kind):
    the reference's own AMRRTplanner::CrashDetection compiled in place (oracle/_ref), else the oracle port."""
    from oracle import loader as orc

    try:
        ref = orc.RefAMRRT(S)
        ref.set_dobs(D)
        return (lambda e: ref.crash_mt(e, threads)), "reference"
    except Exception:
        return (lambda e: orc.edges_check(S, D, e, DIVISOR, 3, nthreads=threads)), "port"


def cpu_edge_rate(S, D, edges, threads, reps=5):
    """-> (edges/s, kind, sample, verdicts): the WHOLE step's edge list, `reps` times after one warm-up pass, best."""
    fn, kind = cpu_edge_fn(S, D, threads)
    fn(edges[: min(len(edges), 2_000_000)])
    best, verdicts = 0.0, None
    for _ in range(reps):
        t = time.perf_counter()
        verdicts = fn(edges)
        best = max(best, len(edges) / (time.perf_counter() - t))
    return best, kind, f"all {len(edges)} step edges x{reps}, best (persistent thread pool)", verdicts


def c4_edges_host(xy, idx):
    """The step's edge list on the host, in the GPU arm's order: 999 999 tree edges (parent -> child; parent = the nearest
    neighbour with a smaller id, node 0 = root) then the 25e6 rewire edges (node -> its 25-NN, ascending)."""
    n = len(xy)
    ar = np.arange(n)[:, None]
    earlier = (idx < ar) & (idx >= 0)
    first = np.argmax(earlier, axis=1)
    par = np.where(earlier.any(1), idx[ar[:, 0], first], 0)
    par[0] = -1
    tree = np.concatenate([xy[par[1:]], xy[1:]], 1)
    rew = np.concatenate([np.repeat(xy, idx.shape[1], axis=0), xy[idx.reshape(-1)]], 1)
    return np.ascontiguousarray(np.concatenate([tree, rew], 0), np.float32), par.astype(np.int32)


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU CrashDetection over the SAME 25 999 999 step edges as the GPU arm, every
    step, on all host threads (persistent pool, oracle/pool.hpp); rank 0 only."""
    if rank != 0:
        return
    S, (circles, polys) = c4_map(42)
    from oracle import loader as orc

    D = np.zeros_like(S)  # DrawDObsmap with the oracle's OpenCV-pinned rasteriser
    for cx, cy, r in circles:
        orc.raster_circle(D, cx, cy, r)
    for p in polys:
        orc.raster_fillpoly(D, p)
    xy = c4_nodes(S, 1000)
    threads = os.cpu_count() or 1
    # the same edge population as the GPU arm, rebuilt on the CPU (the oracle's exact k-NN gives the same lists)
    idx, _, _ = orc.knn(xy, xy.astype(np.float64), KNN, grid=(MAP_SIZE, MAP_SIZE, 16), nthreads=threads)
    edges, _ = c4_edges_host(xy, idx)
    del idx
    fn, kind = cpu_edge_fn(S, D, threads)
    n = len(edges)
    for _ in range(args.warmup):
        fn(edges)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn(edges)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": SCALING,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(n),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                             "sample": f"all {n} step edges per step (persistent thread pool, 8192-edge chunks)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def workload_config(n_edges):
    return {"workload": "C4: AM-RRT* on synthetic 4096x4096 map, 1e6 tree nodes; step = collision pass over "
                        "999999 tree edges + 25e6 rewire (node -> 25-NN) edges, divisor 2, planes S|D",
            "map": "4096x4096 seed 42", "nodes": N_NODES, "k": KNN, "divisor": DIVISOR, "edges_per_step": int(n_edges),
            "l2_policy": "inputs larger than L2 (about 250 MB of plan + tree read per step; the explicit-list extra reads 416 MB)"}


# ------------------------------------------------------------------------------------ main
_RESULT_FD = None


def emit(line):
    """The one JSON line, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    # stdout carries the one JSON line and nothing else: native libraries (NCCL prints its version with printf) write to
    # file descriptor 1, so point that at stderr for the run and keep the real stdout for the result line
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return 0

    import ctypes as C

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # every rank builds the identical workload (seed 0): for N > 1 the tree is REPLICATED and one tick is split over the ranks
    wl = Workload(local_rank, seed=0)
    ctx = wl.ctx
    peer = None
    if world > 1:
        from rtpp_b200 import shard

        peer = shard.PeerTick(ctx, DIVISOR)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    last = [None]

    def step():
        if peer is None:
            wl.step()
        else:
            last[0] = peer.check(3)  # this rank's share; the verdict words of the WHOLE tick land in every rank's buffers

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    launches0 = ctx.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host0 = time.perf_counter()
    ev0.record(wl.stream)
    for _ in range(args.steps):
        step()
    ev1.record(wl.stream)
    barrier()
    t_host1 = time.perf_counter()
    ms_total = ev0.elapsed_time(ev1)
    launches = ctx.kernel_launches() - launches0
    if peer is not None:
        peer.status()  # raises if an exchange timed out (the buffers would read "blocked")
    # keep the GPU busy a little longer if the timed region was too short for NVML's sampling period
    if t_host1 - t_host0 < 0.25:
        t_b0 = time.perf_counter()
        while time.perf_counter() - t_b0 < 0.3:
            for _ in range(50):
                step()
            barrier()
        t_host0, t_host1 = t_b0, time.perf_counter()
    sampler.stop_flag = True
    sampler.join(timeout=1.0)
    clocks = sampler.summary(t_host0, t_host1)

    # max over ranks; for N > 1 the ranks together did ONE step per iteration
    t = torch.tensor([ms_total], dtype=torch.float64, device=wl.dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = wl.n_edges / (ms_step * 1e-3)

    # the timed path's verdicts, for the parity checks below (every rank holds the whole tick)
    if peer is None:
        torch.cuda.synchronize()
        gpu_tree, gpu_rows = wl.tree_bits.cpu().numpy(), wl.row_bits.cpu().numpy()
    else:
        gpu_tree, gpu_rows = peer.read(last[0])
    gpu_verdicts = wl.tick_verdicts(gpu_tree, gpu_rows)

    # ---- e2e: one replanning tick through the host-buffer C-ABI calls.  What changes per tick crosses the boundary inside
    # the timed region: the obstacle poses go in (DrawDObsmap runs on the device), the verdict words come back to page-locked
    # host memory; the tree itself (nodes, parents, k-NN table, plan) is resident, exactly as the planner classes keep it.
    e2e_steps = max(3, min(args.steps, 20))
    h_tree = torch.empty(wl.tree_bits.numel(), dtype=torch.int32, pin_memory=True)
    h_rows = torch.empty(wl.row_bits.numel(), dtype=torch.int32, pin_memory=True)
    circ = np.ascontiguousarray(np.array(wl.obstacles[0], np.int32))
    pxy = np.ascontiguousarray(np.concatenate(wl.obstacles[1]).astype(np.int32))
    psz = np.ascontiguousarray(np.array([len(p) for p in wl.obstacles[1]], np.int32))

    def e2e_step(read=True):
        # MainTask's DrawDObsmap (obstacle poses in) + the tick's collision pass (verdict words out)
        ctx.call("map_draw_dynamic", C.c_void_p(circ.ctypes.data), len(circ), C.c_void_p(pxy.ctypes.data), C.c_void_p(psz.ctypes.data), len(psz))
        if peer is None:
            ctx.call("tick_check", DIVISOR, 3, C.c_void_p(h_tree.data_ptr()), C.c_void_p(h_rows.data_ptr()))  # synchronous
        else:
            d_tree, d_rows = peer.check(3)
            if read:
                ctx.call("d2h", C.c_void_p(h_tree.data_ptr()), C.c_void_p(d_tree), C.c_size_t(h_tree.numel() * 4))
                ctx.call("d2h", C.c_void_p(h_rows.data_ptr()), C.c_void_p(d_rows), C.c_size_t(h_rows.numel() * 4))
            ctx.sync()

    def e2e_time(readers):
        """seconds per tick, max over ranks; `readers` = the ranks that copy the whole tick's verdict words to their host"""
        read = rank in readers
        for _ in range(3):
            e2e_step(read)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step(read)  # returns after the verdict words are back on the host (readers) / the launch has retired
        barrier()
        t = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=wl.dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            peer.status()
        return float(t.item())

    # N > 1: ONE planner process (rank 0) owns the result — it reads the whole tick's verdict words back; the other ranks are
    # workers (poses in, their share of the tick).  The variant in which EVERY rank reads everything back (N copies of 4.3 MB
    # through the box's PCIe complex at once) is timed next to it.
    every_rank_s = e2e_time(set(range(world))) if world > 1 else None
    e2e_s = e2e_time({0})
    same = bool(np.array_equal(h_tree.numpy(), gpu_tree.view(np.int32)) and np.array_equal(h_rows.numpy(), gpu_rows.view(np.int32)))
    e2e = {"value": wl.n_edges / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(circ.nbytes + pxy.nbytes + psz.nbytes),
           "d2h_bytes_per_step": int((h_tree.numel() + h_rows.numel()) * 4), "steps": e2e_steps, "matches_resident_run": same,
           "api": ("rtpp_map_draw_dynamic + rtpp_tick_check (host buffers): obstacle poses in, verdict words out (page-locked, written by the kernel), tree resident"
                   if peer is None else
                   "every rank: rtpp_map_draw_dynamic + rtpp_tick_check_allgather_dev; rank 0 (the planner process) also copies the whole tick's "
                   "verdict words to page-locked host memory; max over ranks")}
    if every_rank_s:
        e2e["every_rank_reads_back"] = {"value": wl.n_edges / every_rank_s, "note": "all ranks copy the whole result to their own host at the same time"}
    # the same edges as an explicit end-point list from the host (rtpp_edges_check): PCIe-bound by 16 B/edge
    h_edges = torch.empty((wl.n_edges, 4), dtype=torch.float32, pin_memory=True)
    h_edges.copy_(wl.edges)
    if rank == 0:
        h_bits2 = torch.empty(wl.bits.numel(), dtype=torch.int32, pin_memory=True)
        for _ in range(2):
            ctx.call("edges_check", C.c_void_p(h_edges.data_ptr()), C.c_size_t(wl.n_edges), DIVISOR, 3, C.c_void_p(h_bits2.data_ptr()))
        t0 = time.perf_counter()
        for _ in range(5):
            ctx.call("edges_check", C.c_void_p(h_edges.data_ptr()), C.c_size_t(wl.n_edges), DIVISOR, 3, C.c_void_p(h_bits2.data_ptr()))
        e2e["explicit_edge_list"] = {"value": wl.n_edges / ((time.perf_counter() - t0) / 5), "h2d_bytes_per_step": wl.n_edges * 16,
                                     "d2h_bytes_per_step": int(h_bits2.numel() * 4), "api": "rtpp_edges_check (host buffers, page-locked), one GPU"}

    # ---- BASELINE configs[4]: the scenario sweep, sharded by instance id over the ranks (all ranks take part)
    sweep = None
    sharded_batch = None
    if not args.no_extras:
        sweep = measure_sweep(ctx, rank, world, torch, dist, wl.dev)
        if world > 1:
            sharded_batch = measure_sharded_batch(wl, rank, world, torch, dist)

    if rank != 0:
        if world > 1:
            dist.barrier()
            peer.close()
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant (only) kernel of the step
    from oracle import loader as orc

    edges_host = h_edges.numpy()
    sample = edges_host[:: max(1, wl.n_edges // 2_000_000)]
    probes_per_edge = orc.edges_nprobes_sum(sample, DIVISOR) / len(sample)
    alg_bytes = wl.n_edges * (16 + 4 * probes_per_edge + 0.125)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes / (ms_step * 1e-3) / 1e9
    prof = {}
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "r02_k1w_ncu_summary.json")))
    except Exception:
        pass
    traffic = prof.get("dram_bytes_per_launch")
    # what the launch has to read: neighbour coordinates (8 B/edge slot), slot tables, node table + parents, plan masks
    n_slots = wl.plan["n_slots"]
    moved_model = n_slots * KNN * 8 + n_slots * (8 + 4 + 4 + 4) + N_NODES * (8 + 8) + n_slots * 4 + (N_NODES // 8)
    roofline = {"bound": "hbm", "kernel": "tick_check_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs, burst)" if peaks else "fallback 6650",
                "algorithmic_bytes_per_edge": 16 + 4 * probes_per_edge + 0.125, "mean_probes_per_edge": probes_per_edge,
                "note": "the algorithmic formula (SURVEY 8(d)) charges 16 B of end points per edge and one 4-byte plane word per probe to HBM; "
                        "the windowed kernel streams 8 B of neighbour coordinates per edge and probes a shared-memory copy of the map window, "
                        "so it moves fewer bytes than the formula counts and frac can exceed 1: physical_frac and issue_frac say what bounds it",
                "moved_bytes_model": int(moved_model), "physical_GBs": (traffic or moved_model) / (ms_step * 1e-3) / 1e9,
                "physical_frac": (traffic or moved_model) / (ms_step * 1e-3) / 1e9 / peak}
    if prof.get("warp_instructions_per_launch"):
        # instruction-issue form: warp instructions of one launch (ncu) / launch time vs 148 SMs x 4 schedulers x SM clock
        issue_peak = ctx.sm_count() * 4 * (clocks.get("sm_mhz") or 1965.0) * 1e6
        roofline["issue_frac"] = prof["warp_instructions_per_launch"] / (ms_step * 1e-3) / issue_peak
        roofline["warp_instructions_per_32_edges"] = prof["warp_instructions_per_launch"] / (wl.n_edges / 32)
        roofline["profile"] = prof.get("source")
    try:
        mp = ctx.measure_peaks(8 << 20, 5)
        roofline["same_run_peaks"] = mp
    except Exception as exc:  # the bench line must survive a failed microbenchmark
        roofline["same_run_peaks"] = {"error": str(exc)}

    # ---- CPU baseline on this box's host cores: the WHOLE step, and full parity of the timed GPU path against it
    threads = os.cpu_count() or 1
    cpu_rate, kind, sample_desc, cpu_verdicts = cpu_edge_rate(wl.S, wl.D, edges_host, threads)
    parity_ok = bool(np.array_equal(gpu_verdicts, cpu_verdicts))
    cpu_baseline = {"value": cpu_rate, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample_desc,
                    "verdicts_match_gpu": parity_ok, "verdicts_compared": int(len(cpu_verdicts))}

    extras = {}
    if not args.no_extras:
        l2g = roofline.get("same_run_peaks", {}).get("l2_sector_gather_gbs") if isinstance(roofline.get("same_run_peaks"), dict) else None
        extras = measure_extras(wl, torch, sweep, roofline_formula=(16 + 4 * probes_per_edge + 0.125), peak=peak, cpu_verdicts=cpu_verdicts,
                                l2_gather_gbs=l2g)
        if sharded_batch:
            extras["sharded_edge_batch"] = sharded_batch

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": SCALING, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(wl.n_edges), "clocks": clocks, "e2e": e2e,
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline, "extras": extras}
    if world > 1:
        line["config"]["sharding"] = ("one tick split over the ranks: tree replicated, work items round-robin, verdict words stored into every rank's "
                                      "buffers over NVLink by the kernel (rtpp_tick_check_allgather_dev)")
    emit(line)
    if world > 1:
        dist.barrier()
        peer.close()
        dist.destroy_process_group()
    return 0


def measure_extras(wl, torch, sweep, roofline_formula=None, peak=None, cpu_verdicts=None, l2_gather_gbs=None):
    """Neighbour queries/s and tick latencies on the resident C4 state (outside the timed region)."""
    ctx = wl.ctx
    out = {}

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(wl.stream)
        for _ in range(reps):
            fn()
        e1.record(wl.stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    n = len(wl.xy)
    # the same step as an explicit end-point list (K1 of csrc/edges.cu: streams 16 B per edge, probes the plane through L1/L2)
    ms = timed(wl.explicit_step, 20)
    got = np.unpackbits(wl.bits.cpu().numpy().view(np.uint8), bitorder="little")[: wl.n_edges]
    out["explicit_list_kernel"] = {"kernel": "edge_check_f32_kernel", "ms_per_step": ms, "edges_per_s": wl.n_edges / (ms * 1e-3),
                                   "verdicts_match_cpu": None if cpu_verdicts is None else bool(np.array_equal(got, cpu_verdicts))}
    if roofline_formula and peak:
        out["explicit_list_kernel"]["roofline_frac"] = wl.n_edges * roofline_formula / (ms * 1e-3) / 1e9 / peak
    ms = timed(lambda: ctx.knn_nodes_dev(KNN, wl.knn_idx.data_ptr()), 5)
    out["knn_queries_per_s"] = n / (ms * 1e-3)
    out["knn_all_nodes_ms"] = ms
    out["knn_algorithmic_GBs"] = n * (16 + KNN * 20) / (ms * 1e-3) / 1e9  # 516 B per 25-NN query
    try:  # the CPU side of the same queries (report both forms, state which), ids compared with the GPU's table
        out["knn_cpu"] = measure_knn_cpu(wl.xy, wl.knn_idx.cpu().numpy().reshape(n, KNN))
    except Exception as exc:  # an extra must not cost the bench line
        out["knn_cpu"] = {"error": repr(exc)}
    # the same store queried at 1e6 uniform random points (binary64, as a planner's samples are: no shared cell between
    # consecutive queries), ids + squared distances out; device buffers
    import ctypes as C

    nq = 1_000_000
    q = torch.rand((nq, 2), dtype=torch.float64, device=wl.dev) * MAP_SIZE
    qi = torch.empty((nq, KNN), dtype=torch.int32, device=wl.dev)
    qd = torch.empty((nq, KNN), dtype=torch.float64, device=wl.dev)
    qc = torch.empty(nq, dtype=torch.int32, device=wl.dev)
    ms = timed(lambda: ctx.call("knn_dev", C.c_void_p(q.data_ptr()), C.c_size_t(nq), KNN, C.c_void_p(qi.data_ptr()), C.c_void_p(qd.data_ptr()),
                                C.c_void_p(qc.data_ptr())), 5)
    out["knn_random_queries_per_s"] = nq / (ms * 1e-3)
    out["knn_random_queries_ms"] = ms
    del q, qi, qd, qc
    out["edge_kernels"] = measure_edge_kernels(wl, torch, timed)
    cost0 = np.full(n, 3.4e38, np.float32)
    col0 = np.ones(n, np.uint8)
    cost0[0], col0[0] = 0.0, 0
    ctx.tree_set(wl.parent, cost0, col0)
    ms = timed(lambda: ctx.tree_propagate_dev(0, DIVISOR, 3), 5)
    out["propagate_1e6_tree_edges_ms"] = ms
    out["propagate_edges_per_s"] = (n - 1) / (ms * 1e-3)
    if roofline_formula and peak:
        # bytes(edge) = 40 (ids, coordinates, two cost|parent words) + the edge-check bytes (the step's mean is used for
        # both); these kernels are latency / atomic bound by the survey's own analysis, the fraction says how far from streaming
        out["propagate_algorithmic_GBs"] = (n - 1) * (40 + roofline_formula) / (ms * 1e-3) / 1e9
        out["propagate_roofline_frac"] = out["propagate_algorithmic_GBs"] / peak
    best = None
    for _ in range(3):  # the first call also allocates the 100 MB edge-weight scratch; same start state every time
        ctx.tree_set(wl.parent, cost0, col0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sweeps = ctx.relax_dev(wl.knn_idx.data_ptr(), KNN, DIVISOR, 3, free_only=True)  # synchronous: returns the bucket count
        ms = (time.perf_counter() - t0) * 1e3
        best = ms if best is None else min(best, ms)
    out["rewire_to_quiescence_ms"] = best
    if roofline_formula and peak:
        out["rewire_algorithmic_GBs"] = n * KNN * (40 + roofline_formula) / (best * 1e-3) / 1e9  # every k-NN edge weighed and relaxed at least once
        out["rewire_roofline_frac"] = out["rewire_algorithmic_GBs"] / peak
    out["rewire_sweeps"] = sweeps
    out["rewire_schedule"] = "work list by cost bucket (near-far), blocks chase their improvements through shared-memory queues; wall time of the host call, best of 3"

    def leg(key, fn):  # an extra must not cost the bench line: a failing leg reports its error under its own key
        try:
            out.update(fn())
        except Exception as exc:
            out[key] = {"error": repr(exc)}

    leg("replan", lambda: measure_replanning(ctx, sweep))
    leg("dropin_rtrrt_class", lambda: {"dropin_rtrrt_class": measure_dropin_class()})
    leg("rtfmt_1e5", lambda: measure_fmt(ctx))
    leg("rtfmt_sweep", lambda: measure_fmt_sweep(ctx))
    leg("cdrt", lambda: measure_cdrt(ctx))
    leg("amrrt", lambda: measure_amrrt(ctx))
    # last: its edge batches load GAME0 into the ctx's planes (nothing after it uses the C4 planes)
    leg("c1_rtrrt_game0", lambda: measure_c1(ctx, torch, timed, peak or 6650.0, l2_gather_gbs))
    return out


def measure_knn_cpu(xy, gpu_idx):
    """Neighbour queries on the host cores over the C4 node set (every 5th node queries its 25-NN): the oracle's exact binary64
    k-NN with its grid accelerator on all threads — the fair CPU form of what the kernels do — and the brute-force definition on
    one thread (what the compiled reference runs on here in place of CGAL's kd-tree, which it rebuilds after every insert).  Both
    id lists are compared with the rows of the GPU's table."""
    from oracle import loader as orc

    threads = os.cpu_count() or 1
    ids = np.arange(0, len(xy), 5)
    t0 = time.perf_counter()
    gi, _, _ = orc.knn(xy, xy[ids].astype(np.float64), KNN, grid=(MAP_SIZE, MAP_SIZE, 16), nthreads=threads)
    dt = time.perf_counter() - t0
    qb = ids[:64]
    t0 = time.perf_counter()
    bi, _, _ = orc.knn(xy, xy[qb].astype(np.float64), KNN, nthreads=1)
    dtb = time.perf_counter() - t0
    return {"grid_queries_per_s": len(ids) / dt, "threads": threads, "kind": "port (oracle grid k-NN, exact binary64 (d2, id) order)",
            "sample": f"{len(ids)} of the {len(xy)} node queries", "ids_match_gpu": bool(np.array_equal(gi, gpu_idx[ids])),
            "brute_force_queries_per_s": len(qb) / dtb, "brute_force_threads": 1, "brute_force_sample": f"{len(qb)} queries",
            "brute_force_ids_match_gpu": bool(np.array_equal(bi, gpu_idx[qb]))}


def measure_dropin_class(ticks=24, dt=5000):
    """BASELINE configs[0] through the drop-in boundary itself: the C++ RTRRTplanner class of rtpp_b200/planners (same method
    names as the reference's, GPU engine underneath) against the reference's OWN class compiled in place (oracle/_ref: RTRRT.cpp
    unmodified; its two absent third-party libraries are stand-ins — cv::Mat shim, brute-force exact k-NN for CGAL), both driven
    through DrawDObsmap + UpdateTree + ExpandRewiring(50 ms of virtual time = 1e4 clock reads) + GetPath on bundled GAME0 with the
    scripted obstacle set of RTRRT/test/testD.cpp, same rand() stream; trees compared bit for bit at the end."""
    from oracle import loader as orc
    from rtpp_b200 import capi

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import TESTD_OUTLINES

    S = orc.load_map("GAME0")
    rand = orc.rand31(5, 3_000_000)
    obs = [(0, 370, 388, 0.0, 57, [(0, 0)]), (0, 2370, 630, 0.0, 114, [(0, 0)]),
           (1, 1364, 1364, 0.78539816, 0, TESTD_OUTLINES[3]), (1, 1340, 2520, 0.0, 0, TESTD_OUTLINES[4])]
    out = {"map": "GAME0", "ticks": ticks, "virtual_dt_ns": dt}
    trees = []
    for name in ("gpu_class", "reference_class"):
        try:
            pl = capi.PlannerRTRRT(S) if name == "gpu_class" else orc.RefRTRRT(S)
        except FileNotFoundError:
            out[name] = "unavailable (oracle/_ref not built)"
            continue
        pl.set_clock(dt)
        pl.set_rand(rand)
        pl.set_agent(350.0, 2350.0)
        pl.set_goal(1400.0, 2050.0)
        for num, (typ, x, y, ang, rad, outline) in enumerate(obs):
            pl.add_obstacle(num, typ, x, y, ang, rad, outline)
        lat = []
        for _ in range(ticks):
            t0 = time.perf_counter()
            if name == "gpu_class":
                pl.main_task()
            else:
                pl.draw_dobs(), pl.update_tree(), pl.expand_rewiring(50)
            path = pl.get_path()
            lat.append(time.perf_counter() - t0)
        trees.append(pl.dump())
        out[name + "_tick_ms_p50"] = float(np.median(lat)) * 1e3
        out[name + "_nodes"] = int(len(trees[-1][0]))
        out[name + "_path_len"] = int(len(path[0]))
        pl.close()
    if len(trees) == 2:
        out["trees_match"] = bool(all(np.array_equal(a.view(np.uint8), b.view(np.uint8)) for a, b in zip(trees[0], trees[1])))
    out["reference_class_note"] = ("the reference class runs on stand-ins for its two absent libraries: brute-force exact k-NN in place of CGAL's "
                                   "kd-tree (rebuilt on the first query after every insert in the original) and a cv::Mat shim; at the ~500 nodes of this "
                                   "scenario the stand-in k-NN is not slower than a rebuilt kd-tree would be, but this is not a CGAL timing")
    return out


def measure_edge_kernels(wl, torch, timed):
    """The other edge routines of SURVEY 8(a) on the same resident C4 edge list (device pointers, CUDA events): E1 (divisor 5, both
    planes: RT-RRT*, RT-FMT), E2 (static plane only), E4 (first blocked point of a linear march, AM-RRT* Steer) and E5b (binary64
    end points, divisor 10, dynamic plane: CDRT-RRT*) next to the timed E3 step."""
    import ctypes as C

    ctx, n = wl.ctx, wl.n_edges
    bits = torch.zeros_like(wl.bits)
    out = {}
    for name, div, mask in (("E1_div5_SD", 5, 3), ("E2_div5_S", 5, 1)):
        ms = timed(lambda: ctx.edges_check_dev(wl.edges.data_ptr(), n, div, mask, bits.data_ptr()), 10)
        out[name] = {"edges": n, "ms": ms, "edges_per_s": n / (ms * 1e-3)}
    m = 4_000_000  # first-hit also writes 8 B per edge; binary64 edges are 32 B each
    sub = wl.edges[wl.n_edges - m:].contiguous()
    pstart = torch.empty((m, 2), dtype=torch.float32, device=wl.dev)
    ms = timed(lambda: ctx.call("edges_first_hit_dev", C.c_void_p(sub.data_ptr()), C.c_size_t(m), C.c_void_p(bits.data_ptr()), C.c_void_p(pstart.data_ptr())), 10)
    out["E4_first_hit"] = {"edges": m, "ms": ms, "edges_per_s": m / (ms * 1e-3)}
    sub64 = sub.double().contiguous()
    ms = timed(lambda: ctx.call("edges_check_f64_dev", C.c_void_p(sub64.data_ptr()), C.c_size_t(m), 10, 2, C.c_void_p(bits.data_ptr())), 10)
    out["E5b_f64_div10_D"] = {"edges": m, "ms": ms, "edges_per_s": m / (ms * 1e-3)}
    return out


def measure_amrrt(ctx, n_inst=1024, ticks=4, dt=20000, cap_rand=60_000):
    """AM-RRT* MainTask (UpdateTree + {Expand with Steer and the diffusion metric, RewireRoot} for 80 ms of virtual time + GetPath)
    on bundled MESS6 with a synthetic diffusion table: 1024 planners (own rand() streams) on the engine against the CPU port on
    the host threads, then one planner alone (tick latency); trees compared bit for bit."""
    from concurrent.futures import ThreadPoolExecutor

    import rtpp_b200
    from oracle import loader as orc

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import synthetic_diffusion

    S = orc.load_map("MESS6")
    H, W = S.shape
    cells, mat, ratio = synthetic_diffusion(W)
    goals = [(70.0, 70.0), (1240.0, 70.0), (1240.0, 1240.0), (70.0, 1240.0)]  # RTRRT/test/testD.cpp:47
    streams = [orc.rand31(500 + i, cap_rand * ticks) for i in range(n_inst)]

    def gpu(n):
        eng = rtpp_b200.Engine(ctx, n, 4096, 5, W, H, cap_rand, kind=2)
        eng.set_map(0, 1, S)
        eng.set_diffusion(cells, mat, ratio)
        agent = np.tile(np.array([[720.0, 720.0]], np.float32), (n, 1))
        goal = np.array([goals[i % 4] for i in range(n)], np.float32)
        eng.init_instances(np.zeros(n, np.int32), agent, goal)
        pos = np.zeros(n, np.int64)
        lat = []
        for _ in range(ticks):
            chunk = np.stack([streams[i][pos[i]: pos[i] + cap_rand] for i in range(n)])
            eng.set_am_inputs(rand=chunk, n_rand=np.full(n, cap_rand, np.int32), outpath_nonempty=np.zeros(n, np.int32))
            ctx.sync()
            t0 = time.perf_counter()
            eng.tick(update_tree=True, run_ms=80, dt_ns=dt, get_path=True)
            ctx.sync()
            lat.append(time.perf_counter() - t0)
            st = eng.stats()
            pos += st[:, 14]
        return eng, lat, st

    eng, lat, st = gpu(n_inst)
    iters_gpu, status = int(st[:, 4].sum()), int(st[:, 9].max())
    threads = os.cpu_count() or 1
    m = min(n_inst, threads)

    def cpu(i):
        o = orc.OrcAMRRT(S)
        o.set_diffusion(cells, mat, ratio)
        o.set_rand(streams[i])
        o.set_clock(dt, 0)
        o.set_agent(720.0, 720.0)
        o.set_goal(*goals[i % 4])
        tl = []
        for _ in range(ticks):
            t0 = time.perf_counter()
            o.main_task()
            o.get_path()
            tl.append(time.perf_counter() - t0)
        d = o.dump()
        o.close()
        return d, tl

    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        res = list(ex.map(cpu, range(m)))
    cpu_s = time.perf_counter() - t0
    same = all(all(np.array_equal(a.view(np.uint8), b.view(np.uint8)) for a, b in zip(res[i][0], eng.read_tree(i))) for i in range(m))
    eng.close()
    _, cpu_alone = cpu(0)  # one planner on an otherwise idle host
    eng1, lat1, _ = gpu(1)
    eng1.close()
    return {"amrrt": {"map": "MESS6", "instances": n_inst, "ticks": ticks, "virtual_dt_ns": dt, "gpu_status_max": status,
                      "gpu_sweep_tick_ms_p50": float(np.median(lat) * 1e3), "gpu_planner_ticks_per_s": n_inst * ticks / sum(lat),
                      "gpu_expand_iters_per_s": iters_gpu / sum(lat), "cpu_instances": m, "cpu_threads": threads,
                      "cpu_planner_ticks_per_s": m * ticks / cpu_s, "cpu_kind": "port (oracle loops)", "trees_match_cpu": bool(same),
                      "gpu_single_instance_tick_ms_p50": float(np.median(lat1) * 1e3),
                      "cpu_single_thread_tick_ms_p50": float(np.median(cpu_alone) * 1e3)}}


def measure_cdrt(ctx):
    """BASELINE configs[1]: CDRT-RRT* on bundled GAME0 with its four moving obstacles, continuous replanning on
    virtual time (tools/cdrt_bench.py): tick latency of one planner over the scripted loop's 600 ticks (30 s at 20 Hz,
    CDRT2/test/testD.cpp:197-241) and the throughput of 1024 planners (one warp each) against the CPU port on the host
    threads; states compared bit for bit."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import cdrt_bench
    from oracle import loader as orc

    return {"cdrt": {"single": cdrt_bench.single(ctx, orc, ticks=600), "sweep": cdrt_bench.sweep(ctx, orc, n_inst=1024, threads=os.cpu_count() or 1)}}


def measure_fmt(ctx):
    """BASELINE configs[2] at its stated size: RT-FMT on GAME0 with 1e5 free-space samples (the reference ships 2000), rounds of
    ExpandFMT + both rewires until V_unv is exhausted (tools/config_runs.py c3_rtfmt); the CPU port (1 thread: one planner
    instance is sequential) runs the first 40 000 rounds and its state is compared with the engine's after the same rounds
    (the run to exhaustion is compared in tests/test_configs_gpu.py)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import config_runs as cr
    from oracle import loader as orc

    return {"rtfmt_1e5": cr.c3_rtfmt(ctx, orc, check_rounds=40_000)}


def measure_c1(ctx, torch, timed, peak, l2_gather_gbs=None):
    """BASELINE configs[0] at its stated size: RT-RRT* on GAME0 with 20 000 injected samples (tree compared with
    the CPU port after the first 3000 ticks; the full run is compared in tests/test_configs_gpu.py), then the 2^24 (node, one of
    its 25-NN) edge batch and the 2^20 long-edge batch through rtpp_edges_check_dev (E1: divisor 5, S|D), each with its own
    roofline by SURVEY 8(d)'s formula and every 16th verdict compared with the CPU oracle."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import config_runs as cr
    from oracle import loader as orc

    out, tree, S = cr.c1_rtrrt(ctx, orc, check_ticks=3000)
    batches = cr.c1_edge_batches(ctx, orc, tree[0], S, check=16, timed=lambda fn: timed(fn, 10))
    for r in batches.values():
        r["roofline_frac"] = r["algorithmic_GBs"] / peak
        # SURVEY 8(d)'s second form for batches whose probes dominate: one 32-byte sector per probe against the L2 sector-gather rate
        # measured in this run (rtpp_measure_peaks); at 8.6 (k-NN edges) and 204 (long edges) probes per edge this, not the 16 B of
        # end points, is what the explicit-list kernel is up against
        r["l2_sector_GBs"] = r["edges_per_s"] * r["mean_probes_per_edge"] * 32 / 1e9
        if l2_gather_gbs:
            r["l2_form_frac"] = r["l2_sector_GBs"] / l2_gather_gbs
    out["edge_batches"] = batches
    return {"c1_rtrrt_game0": out}


def measure_fmt_sweep(ctx, n_inst=512, samples=2000, iters=2000, cpu_inst=16):
    """RT-FMT as shipped (2000 samples per planner, RTFMT/test/test.cpp) over many planners: n_inst instances on GAME0 with their
    own sample sets (rand() seeds) and goals run `iters` rounds of ExpandFMT + both rewires + GetPath on the engine (one warp
    each); the CPU port runs cpu_inst of them on the host threads; the trees of those are compared bit for bit."""
    from concurrent.futures import ThreadPoolExecutor

    import rtpp_b200
    from oracle import loader as orc

    S = orc.load_map("GAME0")
    H, W = S.shape
    goals = [(2000.0, 400.0), (1400.0, 2050.0), (888.0, 800.0), (2500.0, 1500.0)]
    threads = os.cpu_count() or 1

    def make(i):
        o = orc.OrcRTFMT(S)
        o.set_agent(360.0, 540.0)
        o.set_goal(*goals[i % 4])
        o.start(orc.rand31(100 + i, 3 * samples + 1000), samples)
        return o

    with ThreadPoolExecutor(threads) as ex:
        planners = list(ex.map(make, range(n_inst)))
    nodes = [o.dump()[0] for o in planners]
    eng = rtpp_b200.FmtEngine(ctx, n_inst, max(len(a) for a in nodes) + 32, 5, W, H)
    eng.set_map(0, 1, S)
    eng.start(np.zeros(n_inst, np.int32), nodes, np.array([goals[i % 4] for i in range(n_inst)], np.float32))
    ctx.sync()
    t0 = time.perf_counter()
    eng.tick(update_tree=True, iters=iters, with_rewire=True, get_path=True)
    ctx.sync()
    gpu_s = time.perf_counter() - t0

    def run_cpu(o):
        o.update_tree()
        o.expand(iters, True)
        return o.dump()

    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        cpu = list(ex.map(run_cpu, planners[:cpu_inst]))
    cpu_s = time.perf_counter() - t0
    same = True
    for i, (oxy, opar, oc, ocl, ost) in enumerate(cpu):
        par, c, cl, st = eng.read_tree(i)
        same &= bool(np.array_equal(par, opar) and np.array_equal(st, ost) and np.array_equal(c.view(np.uint32), oc.view(np.uint32)))
    status = int(eng.stats()[:, 12].max())
    eng.close()
    for o in planners:
        o.close()
    return {"rtfmt_sweep": {"map": "GAME0", "instances": n_inst, "samples": samples, "rounds": iters, "gpu_status_max": status,
                            "gpu_tick_ms": gpu_s * 1e3, "gpu_rounds_per_s": n_inst * iters / gpu_s, "cpu_instances": cpu_inst, "cpu_threads": threads,
                            "cpu_rounds_per_s": cpu_inst * iters / cpu_s, "cpu_kind": "port (oracle loops)", "trees_match_cpu": same}}


def measure_sharded_batch(wl, rank, world, torch, dist, reps=20):
    """SURVEY 8(e) grain 2: ONE candidate-edge batch (rank 0's 26 M C4 edges, broadcast once, outside the timed region)
    split into contiguous ranges of 32-edge words over the ranks, planes replicated, the verdict words all-gathered
    over NCCL; strong scaling, time = max over ranks per (check + all-gather), result compared with rank 0's
    single-GPU verdicts."""
    from rtpp_b200 import shard

    n_edges = torch.tensor([wl.n_edges], dtype=torch.int64, device=wl.dev)
    dist.broadcast(n_edges, 0)
    ne = int(n_edges.item())
    edges = wl.edges if rank == 0 else torch.empty((ne, 4), dtype=torch.float32, device=wl.dev)
    dist.broadcast(edges, 0)
    w0, nw, e0, ne_local = shard.word_range(ne, rank, world)
    local = torch.zeros(max(nw, 1), dtype=torch.int32, device=wl.dev)
    sub = edges[e0: e0 + ne_local]

    def step():
        wl.ctx.edges_check_dev(sub.data_ptr(), ne_local, DIVISOR, 3, local.data_ptr())
        return shard.all_gather_words(local[:nw], ne)

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        dist.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(wl.stream)
        for _ in range(reps):
            out = fn()
        ev1.record(wl.stream)
        torch.cuda.synchronize()
        t = torch.tensor([ev0.elapsed_time(ev1) / reps], dtype=torch.float64, device=wl.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), out

    ms_nccl, full = timed(step)
    # the same batch with the all-gather fused into the check: verdict words stored into every peer's buffer by the
    # kernel itself (NVLink P2P, rtpp_edges_check_allgather_dev), no collective call on the data path
    peer = shard.PeerEdgeBatch(wl.ctx, ne)
    ms_fused, d_words = timed(lambda: peer.check(edges.data_ptr(), DIVISOR, 3))
    fused_words = peer.read(d_words)
    same = same_fused = None
    if rank == 0:
        wl.explicit_step()  # rank 0's own single-GPU verdicts of the same edges
        torch.cuda.synchronize()
        same = bool((full == wl.bits).all().item())
        same_fused = bool(np.array_equal(fused_words, wl.bits.cpu().numpy()))
    peer.close()
    return {"edges": ne, "n_gpus": world, "scaling": "strong",
            "fused_p2p": {"ms_per_batch": ms_fused, "edges_per_s": ne / (ms_fused * 1e-3), "matches_single_gpu": same_fused,
                          "how": "2048-edge blocks round-robin over the ranks; K1 stores each block's verdict words into every rank's result buffer (CUDA IPC mappings, NVLink) + device-side flag exchange"},
            "nccl": {"ms_per_batch": ms_nccl, "edges_per_s": ne / (ms_nccl * 1e-3), "matches_single_gpu": same,
                     "how": "contiguous word ranges (rank 0's holds the long tree edges); K1 + all_gather_into_tensor of the verdict words, %d bytes per rank" % (4 * (((ne + 31) // 32 + world - 1) // world))}}


def measure_sweep(ctx, rank, world, torch, dist, dev, n_seeds=32, dt=5000, n_samples=2000, max_ticks=200, check_every=128):
    """BASELINE configs[4] as SURVEY 8(d) defines C5: 8192 independent RT-RRT* planner instances = 8 bundled maps x 32 (start,
    goal) pairs (16 from the reference's scripted tables, 16 seeded free pairs) x 32 sample streams; every instance runs
    replanning ticks (UpdateTree + ExpandRewiring(50 ms) + GetPath under the virtual clock: dt ns per clock read => 10^4 clock
    reads per tick) until it has consumed its 2000 samples, after which ExpandRewiring returns at once (RTRRT.cpp:228-232) —
    the sweep ends when the last instance is there.  Sharded by instance id round-robin over the ranks (strong scaling: the
    instance count is fixed); no data-path collective; the fixed-size per-scenario result records are all-gathered at the
    end (NCCL); tick time = max over ranks.  Rank 0 keeps the trees of every `check_every`-th instance (all its own) for the
    tree-byte comparison against the CPU oracle in measure_replanning."""
    import hashlib

    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sweep_bench as sb
    from oracle import loader as orc
    from rtpp_b200 import shard

    maps = {name: orc.load_map(name) for name in sb.SCEN}
    inst = sb.make_instances_c5(maps, n_seeds)
    n_inst = len(inst)
    ids = shard.instance_ids(n_inst, rank, world)
    lat, st, samples, eng = sb.run_gpu(ctx, orc, [inst[i] for i in ids], maps, n_samples, max_ticks, dt, until_exhausted=True, keep=True)
    # every rank must tick the same number of times for the max-over-ranks tick times to line up: pad with zeros
    nt = torch.tensor([len(lat)], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(nt, op=dist.ReduceOp.MAX)
    n_ticks = int(nt.item())
    t = torch.tensor(list(lat) + [0.0] * (n_ticks - len(lat)), dtype=torch.float64, device=dev)
    # record = n_nodes, samples_used, iters, pops_r, pops_s, status, path_len, path_cost_bits
    rec = torch.from_numpy(np.ascontiguousarray(st[:, [0, 2, 4, 5, 6, 9, 10, 12]])).to(dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        rec = shard.all_gather_records(rec, n_inst)
    lat = t.cpu().numpy()
    rec = rec.cpu().numpy()
    tot = float(lat.sum())
    checked = {}
    if rank == 0:
        for k in range(0, n_inst, check_every):
            if k % world == 0:
                checked[k] = (eng.read_tree(k // world), int(st[k // world, 2]))
    eng.close()
    return {"inst": inst, "maps": maps, "records": rec, "n_samples": n_samples, "ticks": n_ticks, "dt": dt, "checked": checked,
            "samples": {k: samples[k // world] for k in checked},
            "out": {"planner": "RT-RRT*", "instances": n_inst, "n_gpus": world, "sharding": "instance id mod n_gpus; records all-gathered (NCCL)" if world > 1 else "single GPU",
                    "definition": "C5 (SURVEY 8(d)): 8 maps x 32 (start, goal) pairs x 32 seeds, every instance ticks until its 2000 samples are consumed",
                    "scaling": "strong", "ticks": n_ticks, "virtual_dt_ns": dt, "sweep_s": tot,
                    "gpu_planner_ticks_per_s": n_inst * n_ticks / tot,
                    "gpu_iters_per_s": float(rec[:, 2].sum() / tot), "gpu_queue_pops_per_s": float((rec[:, 3] + rec[:, 4]).sum() / tot),
                    "gpu_sweep_tick_ms_p50": float(np.median(lat) * 1e3), "gpu_status_max": int(rec[:, 5].max()),
                    "samples_consumed_min": int(rec[:, 1].min()),
                    "nodes_mean": float(rec[:, 0].mean()), "paths_found": int((rec[:, 6] > 0).sum()),
                    "records_sha256": hashlib.sha256(rec.tobytes()).hexdigest()[:16]}}


def measure_replanning(ctx, sweep):
    """Replan iterations/s and tick latency: the sweep's engine numbers (measure_sweep, all ranks) against the CPU oracle port of
    the same loops on all host threads over every 128th instance of the sweep (same maps, start / goal, injected sample streams,
    the same number of ticks) — their TREES (coordinates, parents, cost bits, flags) are compared byte for byte with the
    engine's — plus the one-planner tick latency."""
    from concurrent.futures import ThreadPoolExecutor

    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sweep_bench as sb
    from oracle import loader as orc

    inst, maps, ticks, dt, n_samples = sweep["inst"], sweep["maps"], sweep["ticks"], sweep["dt"], sweep["n_samples"]
    out = {"replan": dict(sweep["out"])}
    threads = os.cpu_count() or 1
    ks = sorted(sweep["checked"])
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        res = list(ex.map(lambda k: sb.run_cpu_instance(orc, maps[inst[k][1]], inst[k][2], inst[k][3], sweep["samples"][k], ticks, dt, want_tree=True), ks))
    wall = time.perf_counter() - t0
    same = True
    for k, (lat, c, tree) in zip(ks, res):
        gpu_tree, gpu_used = sweep["checked"][k]
        same &= c["samples_used"] == gpu_used and len(tree) == len(gpu_tree)
        same &= all(np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8)) for a, b in zip(tree, gpu_tree))
    out["replan"].update({
        "cpu_iters_per_s": sum(r[1]["iters"] for r in res) / wall,
        "cpu_queue_pops_per_s": sum(r[1]["pops_r"] + r[1]["pops_s"] for r in res) / wall,
        "cpu_planner_ticks_per_s": len(ks) * ticks / wall,
        "cpu_threads": threads, "cpu_kind": "port (oracle loops, grid k-NN; the reference rebuilds a kd-tree per insert)",
        "cpu_sample": f"every 128th of the {len(inst)} instances ({len(ks)}), {ticks} ticks each",
        "cpu_single_thread_tick_ms_p50": float(np.median([x for r in res for x in r[0]]) * 1e3),
        "trees_match_gpu": bool(same), "trees_compared": len(ks)})
    lat1, st1, _ = sb.run_gpu(ctx, orc, inst[:1], maps, n_samples, 9, dt)
    out["replan"]["gpu_single_instance_tick_ms_p50"] = float(np.median(lat1) * 1e3)
    return out


if __name__ == "__main__":
    sys.exit(main())

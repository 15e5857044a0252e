"""GPU (one): per-item profile of ONE rank's launch of the multi-GPU tick (rtpp_tick_check_allgather_dev) on the C4 workload.
The other ranks are emulated as in tests/test_tick_gpu.py: their flags are pre-set, so rank 0's kernel does its share and leaves.
    RTPP_TICK_TIMING=1 RTPP_TICK_ITEMS=/tmp/items.bin python tools/tick_items.py WORLD
Prints how long the items of each kind took (far words, tree words, whole groups, group parts), the slowest ones with their
start time, their position in the deal order and the plan's cost estimate.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import bench

    world = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    path = os.environ.setdefault("RTPP_TICK_ITEMS", "/tmp/tick_items.bin")
    os.environ.setdefault("RTPP_TICK_TIMING", "1")
    wl = bench.Workload(0, seed=0)
    ctx = wl.ctx
    info = ctx.tick_plan_info(bench.DIVISOR)
    n = bench.N_NODES
    tree_words, n_slots = (n - 1 + 31) // 32, info["n_slots"]

    def peer_alloc(nbytes):
        p, h = C.c_void_p(), (C.c_ubyte * 64)()
        ctx.call("peer_alloc", C.c_size_t(nbytes), C.byref(p), h)
        return p.value

    bufs = [[peer_alloc(tree_words * 4 + 64), peer_alloc(n_slots * 4)] for _ in range(world)]
    flags = [peer_alloc(64) for _ in range(world)]
    ttab = (C.c_void_p * world)(*[b[0] for b in bufs])
    rtab = (C.c_void_p * world)(*[b[1] for b in bufs])
    ftab = (C.c_void_p * world)(*flags)
    for epoch in (1, 2, 3):
        later = np.zeros(16, np.int32)
        later[1:world] = epoch
        ctx.h2d(flags[0], later)
        ctx.call("tick_check_allgather_dev", bench.DIVISOR, 3, ttab, rtab, ftab, 0, world, epoch)
        ctx.sync()
    raw = np.fromfile(path, np.uint8)
    hdr = raw[:64].view(np.int64)
    n_items, nfw, ntw, stride, first, ncodes, ncost, n_static = [int(v) for v in hdr]
    off = 64
    rec = raw[off: off + 8 * n_items].view(np.uint64)
    off += 8 * n_items
    codes = raw[off: off + 4 * ncodes].view(np.int32)
    off += 4 * ncodes
    cost = raw[off: off + 4 * ncost].view(np.int32)
    dur = (rec & np.uint64(0xffffffff)).astype(np.float64) * 1e-3
    start = (rec >> np.uint64(32)).astype(np.int64)
    seen = dur > 0
    start = (start - start[seen].min()) * 1e-3
    item = np.arange(n_items)
    sh = (item - nfw) * stride + first
    kind = np.where(item < nfw, 0, np.where(sh < ntw, 1, 2))
    gi = np.clip(sh - ntw, 0, max(ncodes - 1, 0))
    code = np.where(kind == 2, codes[gi] if ncodes else gi, 0)
    part = code >> 28
    grp = code & 0x0fffffff
    kind = np.where((kind == 2) & (part > 0), 3, kind)
    names = ["far word", "tree word", "group", "group part"]
    print(f"world {world}: items {n_items} (far {nfw}, static {n_static}), logged {int(seen.sum())}; launch span {start[seen].max() + dur[seen][np.argmax(start[seen])]:.1f} us")
    for kd in range(4):
        m = seen & (kind == kd)
        if m.any():
            d = dur[m]
            print(f"  {names[kd]:10s}: {int(m.sum()):6d} items, mean {d.mean():6.2f} us, p50 {np.median(d):6.2f}, p99 {np.percentile(d, 99):6.2f}, max {d.max():6.2f}, total {d.sum() / 1e3:7.2f} ms")
    top = np.argsort(-dur)[:15]
    for i in top:
        extra = f"group {grp[i]} part {part[i]} cost {cost[grp[i]]}" if kind[i] >= 2 and ncost else ""
        print(f"  item {i:6d} {names[kind[i]]:10s} start +{start[i]:6.1f} us  took {dur[i]:6.2f} us  {'static' if i < n_static else 'drawn '}  {extra}")
    m = seen & (kind >= 2) & (ncost > 0)
    if m.any():
        c = cost[grp[m]].astype(np.float64)
        print("  correlation(duration, cost estimate) over group items: %.3f" % np.corrcoef(dur[m], c)[0, 1])
        late = m & (start > np.percentile(start[m], 50))
        print("  mean duration of group items started in the first / second half of the launch: %.2f / %.2f us" % (dur[m & ~late].mean(), dur[late].mean()))


if __name__ == "__main__":
    main()

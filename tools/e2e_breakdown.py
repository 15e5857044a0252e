"""GPU: where the end-to-end tick (bench.py's e2e) spends its time on the C4 workload: DrawDObsmap on the device, the tick kernel
writing into device memory / page-locked host memory, and the host-side call overhead.  python tools/e2e_breakdown.py"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import bench

    wl = bench.Workload(0, seed=0)
    ctx = wl.ctx
    h_tree = torch.empty(wl.tree_bits.numel(), dtype=torch.int32, pin_memory=True)
    h_rows = torch.empty(wl.row_bits.numel(), dtype=torch.int32, pin_memory=True)
    circ = np.ascontiguousarray(np.array(wl.obstacles[0], np.int32))
    pxy = np.ascontiguousarray(np.concatenate(wl.obstacles[1]).astype(np.int32))
    psz = np.ascontiguousarray(np.array([len(p) for p in wl.obstacles[1]], np.int32))

    def draw():
        ctx.call("map_draw_dynamic", C.c_void_p(circ.ctypes.data), len(circ), C.c_void_p(pxy.ctypes.data), C.c_void_p(psz.ctypes.data), len(psz))

    def wall(fn, reps=50):
        for _ in range(5):
            fn()
        ctx.sync()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        ctx.sync()
        return (time.perf_counter() - t0) / reps * 1e6

    def dev_tick():
        wl.step()
        ctx.sync()

    def host_tick():
        ctx.call("tick_check", bench.DIVISOR, 3, C.c_void_p(h_tree.data_ptr()), C.c_void_p(h_rows.data_ptr()))

    def dev_tick_copy():
        wl.step()
        h_rows.copy_(wl.row_bits, non_blocking=True)
        h_tree.copy_(wl.tree_bits, non_blocking=True)
        ctx.sync()

    out = {"draw_dynamic_us": wall(lambda: (draw(), ctx.sync())), "draw_dynamic_async_us": wall(draw),
           "tick_device_out_sync_us": wall(dev_tick), "tick_device_out_async_us": wall(wl.step),
           "tick_pinned_host_out_us": wall(host_tick), "tick_device_out_then_d2h_us": wall(dev_tick_copy),
           "draw_plus_host_tick_us": wall(lambda: (draw(), host_tick())), "d2h_bytes": int((h_tree.numel() + h_rows.numel()) * 4)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

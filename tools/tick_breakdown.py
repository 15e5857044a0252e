"""Device-time breakdown of the tick-level e2e call (development aid)."""
import os, sys, time, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

wl = bench.Workload(0, 0)
ctx = wl.ctx
def timed(fn, reps=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(wl.stream)
    for _ in range(reps): fn()
    e1.record(wl.stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
bits = torch.zeros_like(wl.bits)
ne = C.c_size_t()
print("explicit-list kernel ms", timed(wl.step))
print("resident-tree kernel ms", timed(lambda: ctx.call("tick_edges_check_dev", bench.DIVISOR, 3, C.c_void_p(bits.data_ptr()), C.byref(ne))))
print("same verdicts", bool((bits == wl.bits).all().item()), ne.value)
d_plane = torch.from_numpy(wl.D).to(wl.dev)
print("pack+merge dynamic plane ms", timed(lambda: ctx.call("map_set_plane_dev", 2, C.c_void_p(d_plane.data_ptr()), bench.MAP_SIZE, bench.MAP_SIZE)))
h_plane = torch.from_numpy(wl.D.copy()).pin_memory()
print("H2D 16.8MB ms", timed(lambda: d_plane.copy_(h_plane, non_blocking=True)))
h_bits = torch.empty_like(bits, device="cpu").pin_memory()
print("D2H 3.25MB ms", timed(lambda: h_bits.copy_(bits, non_blocking=True)))
t0 = time.perf_counter()
for _ in range(20):
    ctx.call("tick_edges_check", C.c_void_p(h_plane.data_ptr()), bench.DIVISOR, 3, C.c_void_p(h_bits.data_ptr()), C.byref(ne))
print("tick call wall ms", (time.perf_counter() - t0) / 20 * 1e3)

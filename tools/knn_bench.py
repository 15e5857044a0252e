#!/usr/bin/env python
"""All-nodes 25-NN table at C4 (1e6 nodes on the synthetic 4096^2 map): kernel time (development aid)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import rtpp_b200  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else bench.KNN
ctx = rtpp_b200.Ctx(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
S, _ = bench.c4_map(42)
xy = bench.c4_nodes(S, 1000)
n = len(xy)
for shift in (4, 5):
    ctx.nodes_reset(shift, n, bench.MAP_SIZE, bench.MAP_SIZE)
    ctx.nodes_append(xy)
    ctx.nodes_build()
    idx = torch.empty((n, k), dtype=torch.int32, device="cuda")
    d2 = torch.empty((n, k), dtype=torch.float64, device="cuda")
    ctx.knn_nodes_dev(k, idx.data_ptr(), d2.data_ptr())
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(5):
        ctx.knn_nodes_dev(k, idx.data_ptr(), d2.data_ptr())
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"cell 2^{shift}: {ms:.3f} ms  {n / ms / 1e3:.1f} M queries/s  checksum {int(idx.long().sum())} {float(d2.sum()):.6f}")

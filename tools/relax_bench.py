"""Bulk relax (K4) on the C4 state: time to quiescence — development aid."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench

def main():
    wl = bench.Workload(0, seed=0)
    ctx = wl.ctx
    n = len(wl.xy)
    cost0 = np.full(n, 3.4e38, np.float32); col0 = np.ones(n, np.uint8); cost0[0], col0[0] = 0.0, 0
    ref = None
    for rep in range(int(os.environ.get("REPS", "3"))):
        ctx.tree_set(wl.parent, cost0, col0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sweeps = ctx.relax_dev(wl.knn_idx.data_ptr(), bench.KNN, bench.DIVISOR, 3, free_only=True)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        par, cost, col = ctx.tree_get()
        h = (int(cost.view(np.uint32).astype(np.uint64).sum()), int(par.astype(np.int64).sum()), int(col.sum()))
        if ref is None: ref = h
        print(json.dumps({"bps": os.environ.get("RTPP_RELAX_BPS"), "ms": round(ms, 3), "rounds": sweeps, "hash": h, "same": h == ref}))

if __name__ == "__main__":
    main()

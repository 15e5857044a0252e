"""Tree propagate (K5) on the C4 state — development aid."""
import os, sys, json
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
    ctx.tree_set(wl.parent, cost0, col0)
    for rep in range(int(os.environ.get("REPS", "3"))):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(wl.stream)
        ctx.tree_propagate_dev(0, bench.DIVISOR, 3)
        e1.record(wl.stream)
        torch.cuda.synchronize()
        par, cost, col = ctx.tree_get()
        print(json.dumps({"ms": round(e0.elapsed_time(e1), 4), "hash": [int(cost.view(np.uint32).astype(np.uint64).sum()), int(col.sum())]}))

if __name__ == "__main__":
    main()

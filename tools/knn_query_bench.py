"""Random-point 25-NN queries at C4 (development aid): uniform over the map vs free space only, unsorted vs cell-sorted."""
import ctypes as C, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench

def main():
    wl = bench.Workload(0, seed=0)
    ctx = wl.ctx
    nq, k = 1_000_000, bench.KNN
    rng = np.random.default_rng(3)
    uni = rng.random((nq, 2)) * bench.MAP_SIZE
    free = uni[wl.S[uni[:, 1].astype(np.int64), uni[:, 0].astype(np.int64)] == 0]
    free = np.concatenate([free, free])[:nq]
    def run(name, pts):
        q = torch.from_numpy(np.ascontiguousarray(pts)).to(wl.dev)
        qi = torch.empty((nq, k), dtype=torch.int32, device=wl.dev); qd = torch.empty((nq, k), dtype=torch.float64, device=wl.dev)
        qc = torch.empty(nq, dtype=torch.int32, device=wl.dev)
        f = lambda: ctx.call("knn_dev", C.c_void_p(q.data_ptr()), C.c_size_t(nq), k, C.c_void_p(qi.data_ptr()), C.c_void_p(qd.data_ptr()), C.c_void_p(qc.data_ptr()))
        f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(wl.stream)
        for _ in range(5): f()
        e1.record(wl.stream); torch.cuda.synchronize()
        print(json.dumps({"case": name, "ms": round(e0.elapsed_time(e1) / 5, 3), "kth_d_mean": float(qd[:, k - 1].sqrt().mean())}))
    def cell_sorted(p):
        key = (p[:, 1].astype(np.int64) >> 4) * (bench.MAP_SIZE >> 4) + (p[:, 0].astype(np.int64) >> 4)
        return p[np.argsort(key, kind="stable")]
    run("uniform", uni); run("uniform cell-sorted", cell_sorted(uni))
    run("free space", free); run("free space cell-sorted", cell_sorted(free))

if __name__ == "__main__":
    main()

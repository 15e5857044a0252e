"""Quick device-side timing of K1 (edge_check) on a resident edge list — development aid."""
import sys, os, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import rtpp_b200
from oracle import loader as orc

def main():
    size = 4096
    S = orc.synthetic_map(size)
    ctx = rtpp_b200.Ctx(0)
    stream = torch.cuda.Stream()  # a non-default stream: its handle is what rtpp_set_stream needs
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.map_set_plane(1, S); ctx.map_clear_plane(2)
    rng = np.random.default_rng(0)
    n = 26_000_000
    for name, spread, order in [("short-random", 10.0, "random"), ("short-sorted", 10.0, "sorted"), ("mix80", 80.0, "random")]:
        a = rng.random((n, 2), dtype=np.float32) * np.float32(size)
        if order == "sorted":
            key = (a[:, 1].astype(np.int32) >> 5) * 128 + (a[:, 0].astype(np.int32) >> 5)
            a = a[np.argsort(key, kind="stable")]
        b = np.clip(a + (rng.random((n, 2), dtype=np.float32) - 0.5) * np.float32(2 * spread), 0, np.float32(size - 0.01))
        e = np.ascontiguousarray(np.concatenate([a, b], 1), np.float32)
        d_e = torch.from_numpy(e).cuda()
        d_b = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
        for div in (2, 5):
            npr = orc.edges_nprobes_sum(e[: 1_000_000], div) / 1_000_000
            for _ in range(3):
                ctx.edges_check_dev(d_e.data_ptr(), n, div, 3, d_b.data_ptr())
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            iters = 10
            for _ in range(iters):
                ctx.edges_check_dev(d_e.data_ptr(), n, div, 3, d_b.data_ptr())
            ev1.record(); torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1) / iters
            bytes_alg = n * (16 + 4 * npr + 0.125)
            print(json.dumps({"case": name, "div": div, "n": n, "ms": round(ms, 4), "Gedges_s": round(n / ms / 1e6, 2),
                              "mean_probes": round(npr, 2), "alg_GBs": round(bytes_alg / ms / 1e6, 1),
                              "stream_GBs": round(n * 16.125 / ms / 1e6, 1)}))
            if name == "short-random" and div == 2:
                got = np.unpackbits(d_b.cpu().numpy().view(np.uint8), bitorder="little")[:n]
                want = orc.edges_check(S, None, e[:2_000_000], div, 1, nthreads=8)
                print("check mismatches:", int((got[:2_000_000] != want).sum()), "blocked", float(want.mean()))
        del d_e, d_b

if __name__ == "__main__":
    main()

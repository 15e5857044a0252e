"""GPU: the windowed tick check (rtpp_tick_check_dev) against the explicit-list K1 on the C4 workload of bench.py:
plan statistics, verdict equality over all 26.0 M edges, and CUDA-event times with and without an L2 flush between launches.
    python tools/tick_bench.py [reps]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import bench

    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    wl = bench.Workload(0, seed=0)
    ctx = wl.ctx
    n, k = bench.N_NODES, bench.KNN
    info = ctx.tick_plan_info(bench.DIVISOR)
    out = {"n_slots": info["n_slots"], "usable": info["usable"], "far_fraction": info["far_fraction"]}
    tree = torch.zeros((n - 1 + 31) // 32, dtype=torch.int32, device=wl.dev)
    rows = torch.zeros(info["n_slots"], dtype=torch.int32, device=wl.dev)
    wl.explicit_step()
    ctx.tick_check_dev(bench.DIVISOR, 3, tree.data_ptr(), rows.data_ptr())
    torch.cuda.synchronize()
    # verdict equality with the explicit list (itself checked against the oracle in tests/)
    bits = np.unpackbits(wl.bits.cpu().numpy().view(np.uint8), bitorder="little")[: wl.n_edges]
    tb = np.unpackbits(tree.cpu().numpy().view(np.uint8), bitorder="little")[: n - 1]
    r = rows.cpu().numpy().view(np.uint32)[info["slot_of_id"]]
    rb = ((r[:, None] >> np.arange(k, dtype=np.uint32)[None, :]) & 1).astype(np.uint8).reshape(-1)
    out["tree_equal"] = bool(np.array_equal(tb, bits[: n - 1]))
    out["rewire_equal"] = bool(np.array_equal(rb, bits[n - 1:]))
    out["blocked_fraction"] = float(bits.mean())
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=wl.dev)
    legacy = torch.zeros_like(wl.bits)

    def timed(fn, with_flush):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(reps):
            if with_flush:
                flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(wl.stream)
            fn()
            e1.record(wl.stream)
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / reps

    import ctypes as C

    ne = C.c_size_t()
    for name, fn in (("explicit_k1", wl.explicit_step),
                     ("tick_window", lambda: ctx.tick_check_dev(bench.DIVISOR, 3, tree.data_ptr(), rows.data_ptr())),
                     ("tick_legacy_order", lambda: ctx.call("tick_edges_check_dev", bench.DIVISOR, 3, C.c_void_p(legacy.data_ptr()), C.byref(ne)))):
        out[name + "_ms_flushed"] = timed(fn, True)
        out[name + "_ms_warm"] = timed(fn, False)
    out["legacy_equal"] = bool(torch.equal(legacy, wl.bits))
    print(json.dumps(out))


if __name__ == "__main__":
    main()

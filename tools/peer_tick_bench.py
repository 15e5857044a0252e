"""GPU (torchrun, N ranks): ONE C4 tick split over the ranks with rtpp_tick_check_allgather_dev (shard.PeerTick): per-rank
CUDA-event time per step, max over ranks, and equality of every rank's gathered verdict words with rank 0's single-GPU tick.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/peer_tick_bench.py [reps]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import bench
    from rtpp_b200 import shard

    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    rank, local, world = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = bench.Workload(local, seed=0)
    peer = shard.PeerTick(wl.ctx, bench.DIVISOR)

    def timed(fn):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(wl.stream)
        for _ in range(reps):
            fn()
        e1.record(wl.stream)
        torch.cuda.synchronize()
        dist.barrier()
        return e0.elapsed_time(e1) / reps

    ms_single = timed(wl.step)
    last = [None]

    def fused():
        last[0] = peer.check(3)

    ms_fused = timed(fused)
    tree, rows = peer.read(last[0])
    wl.step()
    torch.cuda.synchronize()
    same = bool(np.array_equal(tree.view(np.int32), wl.tree_bits.cpu().numpy()) and np.array_equal(rows.view(np.int32), wl.row_bits.cpu().numpy()))
    t = torch.tensor([ms_single, ms_fused, float(same)], dtype=torch.float64, device=wl.dev)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    if rank == 0:
        m = torch.stack(out).cpu().numpy()
        print(json.dumps({"n_gpus": world, "parts": os.environ.get("RTPP_TICK_PARTS", "3"), "single_gpu_ms": m[:, 0].tolist(),
                          "fused_ms_per_rank": m[:, 1].tolist(), "fused_ms": float(m[:, 1].max()), "speedup": float(m[0, 0] / m[:, 1].max()),
                          "all_ranks_match_single_gpu": bool(m[:, 2].all())}))
    peer.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

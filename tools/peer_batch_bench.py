#!/usr/bin/env python
"""Development aid (run under torchrun, 2+ GPUs): pieces of the sharded edge batch — local K1 on the rank's range, the fused
check + P2P all-gather, host-side cost of a call."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import bench
from rtpp_b200 import shard

rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
wl = bench.Workload(local_rank, seed=0)  # same seed everywhere: identical edges on every rank
ne = wl.n_edges
w0, nw, e0, nel = shard.word_range(ne, rank, world)
sub = wl.edges[e0: e0 + nel]
local = torch.zeros(nw, dtype=torch.int32, device=wl.dev)


def timed(fn, reps=50):
    fn()
    torch.cuda.synchronize()
    dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record(wl.stream)
    for _ in range(reps):
        fn()
    ev1.record(wl.stream)
    host = (time.perf_counter() - t0) / reps
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / reps, host * 1e3


peer = shard.PeerEdgeBatch(wl.ctx, ne)
res = {
    "whole batch, one GPU": timed(wl.step),
    "local range only": timed(lambda: wl.ctx.edges_check_dev(sub.data_ptr(), nel, bench.DIVISOR, 3, local.data_ptr())),
    "fused p2p": timed(lambda: peer.check(wl.edges.data_ptr(), bench.DIVISOR, 3)),
    "nccl": timed(lambda: (wl.ctx.edges_check_dev(sub.data_ptr(), nel, bench.DIVISOR, 3, local.data_ptr()), shard.all_gather_words(local, ne))),
}
for r in range(world):
    dist.barrier()
    if r == rank:
        print(f"rank {rank}: edges [{e0}, {e0 + nel})", {k: f"{v[0]:.4f} ms dev, {v[1]:.4f} ms host/call" for k, v in res.items()}, flush=True)
words = peer.read(peer.check(wl.edges.data_ptr(), bench.DIVISOR, 3))
wl.explicit_step()
torch.cuda.synchronize()
print(f"rank {rank}: fused result equals the one-GPU verdicts:", bool((torch.from_numpy(words).to(wl.dev) == wl.bits).all().item()), flush=True)
peer.close()
dist.destroy_process_group()

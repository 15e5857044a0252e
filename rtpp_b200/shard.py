"""Multi-GPU plumbing: how the hot path shards over ranks (one process per GPU).

Only two grains shard; both are embarrassingly parallel:
  1. independent planner instances (map x start/goal x seed): instance ids round-robin over ranks,
     maps replicated, one all-gather of fixed-size result records at the end;
  2. one large candidate-edge batch: contiguous ranges of 32-edge words per rank, planes and nodes
     replicated, one all-gather of the verdict words (1 bit per edge).
k-NN over one tree and the sequential rewire queues do not shard (replicas only).  The collective
is torch.distributed's all_gather (NCCL over NVLink on the GPU box, gloo in the CPU tests): the
payloads are tiny (N/8 bytes; ~100 B per scenario), so latency, not bandwidth, is what matters.
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def instance_ids(n_instances, rank, world_size):
    """Round-robin partition of scenario ids (grain 1)."""
    return list(range(rank, n_instances, world_size))


def word_range(n_edges, rank, world_size):
    """Contiguous range of 32-edge verdict words owned by `rank` (grain 2): returns
    (first_word, n_words, first_edge, n_edges_local).  Ranges are word-aligned so that the
    gathered words concatenate into exactly the single-GPU result."""
    n_words = (n_edges + 31) // 32
    per = (n_words + world_size - 1) // world_size
    w0 = min(n_words, rank * per)
    w1 = min(n_words, w0 + per)
    e0 = w0 * 32
    e1 = min(n_edges, w1 * 32)
    return w0, w1 - w0, e0, max(0, e1 - e0)


def all_gather_words(local_words, n_edges):
    """All-gather the per-rank verdict words (int32 tensor on the collective's device) into the
    full word vector of the batch."""
    rank, ws = world()
    if ws == 1:
        return local_words
    n_words = (n_edges + 31) // 32
    per = (n_words + ws - 1) // ws
    if local_words.numel() == per:
        buf = local_words.contiguous()
    else:  # the last rank's range is shorter: pad to the common length
        buf = torch.zeros(per, dtype=local_words.dtype, device=local_words.device)
        buf[: local_words.numel()] = local_words
    out = torch.empty(ws * per, dtype=local_words.dtype, device=local_words.device)
    try:
        dist.all_gather_into_tensor(out, buf)  # one flat collective (NCCL; recent gloo)
    except (RuntimeError, NotImplementedError):
        parts = [torch.empty_like(buf) for _ in range(ws)]
        dist.all_gather(parts, buf)
        out = torch.cat(parts)
    return out[:n_words]


def all_gather_records(local_records, n_instances):
    """All-gather fixed-size per-instance records (2-D tensor, one row per local instance, rows in
    instance_ids order) and return them ordered by global instance id."""
    rank, ws = world()
    if ws == 1:
        return local_records
    per = (n_instances + ws - 1) // ws
    buf = torch.zeros((per, local_records.shape[1]), dtype=local_records.dtype, device=local_records.device)
    buf[: local_records.shape[0]] = local_records
    out = [torch.empty_like(buf) for _ in range(ws)]
    dist.all_gather(out, buf)
    full = torch.zeros((n_instances, local_records.shape[1]), dtype=local_records.dtype, device=local_records.device)
    for r in range(ws):
        ids = instance_ids(n_instances, r, ws)
        full[ids] = out[r][: len(ids)]
    return full


class PeerEdgeBatch:
    """Grain 2 without a collective on the data path: every rank checks its share of ONE edge batch (2048-edge blocks
    round-robin: edge cost varies with length, contiguous ranges are unbalanced) and its kernel stores the verdict words straight into every rank's result buffer over NVLink (CUDA IPC mappings of
    rtpp_peer_alloc buffers; rtpp_edges_check_allgather_dev), followed by a flag exchange on the device.  torch.distributed
    only carries the 64-byte IPC handles at set-up.  Two result buffers alternate by call parity (include/rtpp_gpu.h)."""

    def __init__(self, ctx, n_edges):
        import ctypes as C

        import numpy as np

        self.C, self.np, self.ctx, self.n_edges = C, np, ctx, n_edges
        self.rank, self.world = world()
        assert dist.is_initialized() and 1 <= self.world <= 8
        self.n_words = (n_edges + 31) // 32
        self.epoch = 0
        mine = []
        handles = []
        for nbytes in (self.n_words * 4, self.n_words * 4, 64):  # two result buffers, then the flags
            ptr, h = C.c_void_p(), (C.c_ubyte * 64)()
            ctx._ck(ctx.L.rtpp_peer_alloc(ctx.h, C.c_size_t(nbytes), C.byref(ptr), h))
            mine.append(ptr.value)
            handles.append(bytes(h))
        everyone = [None] * self.world
        dist.all_gather_object(everyone, handles)
        self.own = mine
        self.opened = []
        self.ptrs = [[0] * self.world for _ in range(3)]  # [buffer kind][rank]
        for r in range(self.world):
            for b in range(3):
                if r == self.rank:
                    self.ptrs[b][r] = mine[b]
                else:
                    p = C.c_void_p()
                    hb = (C.c_ubyte * 64).from_buffer_copy(everyone[r][b])
                    ctx._ck(ctx.L.rtpp_peer_open(ctx.h, hb, C.byref(p)))
                    self.ptrs[b][r] = p.value
                    self.opened.append(p.value)
        self.tabs = [(C.c_void_p * self.world)(*self.ptrs[b]) for b in range(3)]
        dist.barrier()

    def check(self, d_edges, divisor, plane_mask):
        """d_edges: device pointer of the whole batch (replicated on every rank).  Returns the device pointer of the
        whole batch's verdict words (valid once the ctx stream has passed this call)."""
        C = self.C
        self.epoch += 1
        b = self.epoch & 1
        self.ctx._ck(self.ctx.L.rtpp_edges_check_allgather_dev(self.ctx.h, C.c_void_p(d_edges), C.c_size_t(self.n_edges), divisor, plane_mask,
                                                               self.tabs[b], self.tabs[2], self.rank, self.world, self.epoch))
        return self.own[b]

    def read(self, d_words):
        out = self.np.zeros(self.n_words, self.np.int32)
        self.ctx.d2h(out, d_words)
        self.ctx.sync()
        return out

    def close(self):
        self.ctx.sync()
        dist.barrier()
        for p in self.opened:
            self.ctx.L.rtpp_peer_close(self.ctx.h, self.C.c_void_p(p))
        dist.barrier()
        for p in self.own:
            self.ctx.dev_free(p)
        self.opened, self.own = [], []


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


def tick_items_of_rank(n_tree_words, n_groups, rank, world_size):
    """The shared work items of ONE tick that `rank` takes in rtpp_tick_check_allgather_dev (csrc/tick.cu: shared item s is a tree
    word for s < n_tree_words, else position s - n_tree_words of the plan's deal order; rank r takes s = r, r + world, ...).
    Host-side mirror of the kernel's dealing, for tests and for sizing: returns (tree_words, group_positions) as ranges."""
    shared = n_tree_words + n_groups
    mine = range(rank, shared, world_size)
    first_group = next((i for i, s in enumerate(mine) if s >= n_tree_words), len(mine)) if len(mine) else 0
    return mine[:first_group], range(mine[first_group] - n_tree_words, n_groups, world_size) if first_group < len(mine) else range(0)


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


def exchange_peer_handles(digest, handles):
    """Set-up step of the fused multi-GPU calls: every rank contributes (a digest of what must be identical on all ranks —
    the tick's slot order —, its list of 64-byte IPC handles) and gets everybody's, in rank order.  Raises on EVERY rank when
    the digests differ, so that no rank goes on to launch a kernel that would wait for peers working on another plan."""
    rank, ws = world()
    everyone = [None] * ws
    dist.all_gather_object(everyone, (digest, [bytes(h) for h in handles]))
    if any(e[0] != everyone[0][0] for e in everyone):
        raise RuntimeError("the ranks disagree on the tick's slot order (different node sets?)")
    if any(len(e[1]) != len(handles) or any(len(h) != 64 for h in e[1]) for e in everyone):
        raise RuntimeError("malformed peer handle list")
    return everyone


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
        for nbytes in (self.n_words * 4, self.n_words * 4, 64):  # two result buffers, then the flags (RTPP_PEER_FLAG_INTS ints)
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

    def status(self):
        """Synchronises; raises when a fused exchange timed out since the last call (its verdict words read blocked)."""
        t = self.C.c_int()
        self.ctx._ck(self.ctx.L.rtpp_peer_status(self.ctx.h, self.C.c_void_p(self.own[2]), self.C.byref(t)))
        if t.value:
            raise RuntimeError(f"{t.value} fused edge-batch exchange(s) timed out: a peer rank did not arrive; verdicts read blocked")

    def read(self, d_words):
        out = self.np.zeros(self.n_words, self.np.int32)
        self.ctx.d2h(out, d_words)
        self.status()
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



class PeerTick:
    """Grain 2 on the RESIDENT tree: ONE tick (tree edges + node -> k-NN edges of the replicated tree) checked by all ranks of a
    node, rtpp_tick_check_allgather_dev.  The tick's work items go round-robin over the ranks; every warp stores its verdict words
    straight into every rank's result buffers over NVLink (CUDA IPC mappings of rtpp_peer_alloc buffers) and the last block of a
    rank's launch exchanges flags with the peers, so no collective call sits on the data path.  torch.distributed only carries the
    64-byte IPC handles (and a digest of the slot order, which must be the same on every rank) at set-up.  Two result buffer pairs
    alternate by call parity (include/rtpp_gpu.h).  A peer that never arrives makes the call fail closed: status() raises and the
    buffers read "blocked"."""

    FLAG_INTS = 16

    def __init__(self, ctx, divisor):
        import ctypes as C
        import hashlib

        import numpy as np

        self.C, self.np, self.ctx, self.divisor = C, np, ctx, divisor
        self.rank, self.world = world()
        assert dist.is_initialized() and 1 <= self.world <= 8
        info = ctx.tick_plan_info(divisor)
        if not info["usable"]:
            raise RuntimeError("the windowed tick plan does not apply to this node store")
        self.n_slots = info["n_slots"]
        self.slot_of_id = info["slot_of_id"]
        self.tree_words = (max(ctx.nodes_count() - 1, 0) + 31) // 32
        digest = hashlib.sha256(info["slot_id"].tobytes()).hexdigest()
        self.epoch = 0
        mine, handles = [], []
        for nbytes in (self.tree_words * 4, self.n_slots * 4, self.tree_words * 4, self.n_slots * 4, self.FLAG_INTS * 4):
            ptr, h = C.c_void_p(), (C.c_ubyte * 64)()
            ctx._ck(ctx.L.rtpp_peer_alloc(ctx.h, C.c_size_t(max(nbytes, 4)), C.byref(ptr), h))
            mine.append(ptr.value)
            handles.append(bytes(h))
        everyone = exchange_peer_handles(digest, handles)
        self.own, self.opened = mine, []
        self.ptrs = [[0] * self.world for _ in range(5)]  # [buffer kind][rank]: tree0, row0, tree1, row1, flags
        for r in range(self.world):
            for b in range(5):
                if r == self.rank:
                    self.ptrs[b][r] = mine[b]
                else:
                    p = C.c_void_p()
                    hb = (C.c_ubyte * 64).from_buffer_copy(everyone[r][1][b])
                    ctx._ck(ctx.L.rtpp_peer_open(ctx.h, hb, C.byref(p)))
                    self.ptrs[b][r] = p.value
                    self.opened.append(p.value)
        self.tabs = [(C.c_void_p * self.world)(*self.ptrs[b]) for b in range(5)]
        dist.barrier()

    def check(self, plane_mask):
        """Launches this rank's share.  Returns (tree words, row words) device pointers of the WHOLE tick's verdicts on this rank
        (valid once the ctx stream has passed this call)."""
        self.epoch += 1
        b = (self.epoch & 1) * 2
        self.ctx._ck(self.ctx.L.rtpp_tick_check_allgather_dev(self.ctx.h, int(self.divisor), int(plane_mask), self.tabs[b], self.tabs[b + 1], self.tabs[4],
                                                              self.rank, self.world, self.epoch))
        return self.own[b], self.own[b + 1]

    def status(self):
        """Synchronises; raises when a fused exchange timed out since the last call (its verdicts read blocked)."""
        t = self.C.c_int()
        self.ctx._ck(self.ctx.L.rtpp_peer_status(self.ctx.h, self.C.c_void_p(self.own[4]), self.C.byref(t)))
        if t.value:
            raise RuntimeError(f"{t.value} fused tick exchange(s) timed out: a peer rank did not arrive; verdicts read blocked")

    def read(self, ptrs):
        tree = self.np.zeros(max(self.tree_words, 1), self.np.uint32)
        rows = self.np.zeros(max(self.n_slots, 1), self.np.uint32)
        self.ctx.d2h(tree, ptrs[0])
        self.ctx.d2h(rows, ptrs[1])
        self.status()
        return tree[: self.tree_words], rows[: self.n_slots]

    def close(self):
        self.ctx.sync()
        dist.barrier()
        for p in self.opened:
            self.ctx.L.rtpp_peer_close(self.ctx.h, self.C.c_void_p(p))
        dist.barrier()
        for p in self.own:
            self.ctx.dev_free(p)
        self.opened, self.own = [], []

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
    buf = torch.zeros(per, dtype=local_words.dtype, device=local_words.device)
    buf[: local_words.numel()] = local_words
    out = [torch.empty_like(buf) for _ in range(ws)]
    dist.all_gather(out, buf)
    return torch.cat(out)[:n_words]


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

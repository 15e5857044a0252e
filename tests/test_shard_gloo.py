"""CPU, world_size 2 over gloo: the multi-GPU host logic (rtpp_b200/shard.py).  The edge-batch
grain (contiguous word ranges + all-gather of verdict words) and the scenario grain (round-robin
instances + all-gather of result records) reassemble exactly the single-process result.  The
per-rank verdicts come from the CPU oracle here (test infrastructure); on the GPU box the same
plumbing carries the CUDA kernel's words over NCCL."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, ws, port, n_edges, n_inst, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    from helpers import edge_mix
    from oracle import loader as orc
    from rtpp_b200 import shard

    S = orc.load_map("Trap")
    H, W = S.shape
    e = edge_mix(np.random.default_rng(1), n_edges, W, H)
    w0, nw, e0, ne = shard.word_range(n_edges, rank, ws)
    local = orc.edges_check(S, None, e[e0:e0 + ne], 5, 1)
    words = np.packbits(np.pad(local, (0, nw * 32 - ne)), bitorder="little").view(np.int32)
    full = shard.all_gather_words(torch.from_numpy(words.copy()), n_edges)
    ids = shard.instance_ids(n_inst, rank, ws)
    rec = torch.tensor([[i, i * i, rank] for i in ids], dtype=torch.int64).reshape(-1, 3)
    allrec = shard.all_gather_records(rec, n_inst)
    if rank == 0:
        q.put((full.numpy().copy(), allrec.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_edges,n_inst", [(100_003, 11), (64, 2), (31, 1)])
def test_two_rank_gather_matches_single_process(oracle, n_edges, n_inst):
    from helpers import edge_mix

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n_edges) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_edges, n_inst, q)) for r in range(2)]
    for p in procs:
        p.start()
    words, rec = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    S = oracle.load_map("Trap")
    H, W = S.shape
    e = edge_mix(np.random.default_rng(1), n_edges, W, H)
    want = oracle.edges_check(S, None, e, 5, 1)
    got = np.unpackbits(words.view(np.uint8), bitorder="little")[:n_edges]
    assert np.array_equal(got, want)
    assert rec[:, 0].tolist() == list(range(n_inst)) and rec[:, 1].tolist() == [i * i for i in range(n_inst)]
    assert rec[:, 2].tolist() == [i % 2 for i in range(n_inst)]


def test_partitions_cover_everything_once():
    from rtpp_b200 import shard

    for n in (0, 1, 31, 32, 33, 1000, 26_000_000):
        for ws in (1, 2, 4, 8):
            tot, nxt = 0, 0
            for r in range(ws):
                w0, nw, e0, ne = shard.word_range(n, r, ws)
                assert e0 == min(nxt, ((n + 31) // 32) * 32) or ne == 0
                nxt = e0 + ne if ne else nxt
                tot += ne
            assert tot == n
            ids = sorted(i for r in range(ws) for i in shard.instance_ids(n % 1000, r, ws))
            assert ids == list(range(n % 1000))


def test_tick_items_are_dealt_once_over_the_ranks():
    """The dealing of one tick's shared items (tree words, then window groups) over the ranks of rtpp_tick_check_allgather_dev:
    every item goes to exactly one rank and the shares differ by at most one item."""
    from rtpp_b200 import shard

    for ntw, ng in ((0, 0), (1, 0), (0, 5), (31250, 32896), (7, 3), (1000, 1)):
        for ws in (1, 2, 3, 4, 8):
            tree, groups, sizes = [], [], []
            for r in range(ws):
                t, g = shard.tick_items_of_rank(ntw, ng, r, ws)
                tree += list(t)
                groups += list(g)
                sizes.append(len(t) + len(g))
            assert sorted(tree) == list(range(ntw)) and sorted(groups) == list(range(ng))
            assert max(sizes) - min(sizes) <= 1


def _handle_worker(rank, ws, port, same_plan, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    from rtpp_b200 import shard

    handles = [bytes([rank * 16 + b]) * 64 for b in range(5)]
    digest = "plan-A" if (same_plan or rank == 0) else "plan-B"
    try:
        got = shard.exchange_peer_handles(digest, handles)
        q.put((rank, "ok", [[h[0] for h in e[1]] for e in got]))
    except RuntimeError as exc:
        q.put((rank, "raised", str(exc)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("same_plan", [True, False])
def test_peer_handle_exchange_and_plan_agreement(same_plan):
    """Set-up of the fused tick (shard.PeerTick): handles arrive in rank order on every rank; ranks whose plans differ ALL refuse
    (a rank that went on alone would launch a kernel waiting for peers that never come)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 977 + int(same_plan)) % 2000
    procs = [ctx.Process(target=_handle_worker, args=(r, 2, port, same_plan, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    if same_plan:
        want = [[0, 1, 2, 3, 4], [16, 17, 18, 19, 20]]
        assert [r[1] for r in res] == ["ok", "ok"] and res[0][2] == want and res[1][2] == want
    else:
        assert [r[1] for r in res] == ["raised", "raised"] and "slot order" in res[0][2]

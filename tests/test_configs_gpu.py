"""GPU: BASELINE.json's configured workloads at their STATED sizes, each compared with the CPU oracle port of the
same loop on the same injected streams (the oracle reproduces the reference's own compiled code: tests/test_oracle_ref.py):
C1 RT-RRT* on GAME0 with 20 000 samples + the 2^24 / 2^20 edge batches, C2 CDRT-RRT* through 600 scripted ticks, C3 RT-FMT
with 1e5 samples until V_unv is exhausted, C5 the 8 x 32 x 32 scenario sweep.  (C4 = tests/test_c4_full_size_gpu.py.)"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu


def test_c1_rtrrt_20000_samples_and_edge_batches(ctx, oracle):
    import config_runs as cr

    out, tree, S = cr.c1_rtrrt(ctx, oracle)
    assert out["samples_used"] == 20_000 and out["gpu_status"] == 0
    assert out["tree_matches_cpu"], out
    assert out["nodes"] > 5000
    batches = cr.c1_edge_batches(ctx, oracle, tree[0], S)
    for name, r in batches.items():
        assert r["verdicts_match_cpu"] and r["verdicts_compared"] == r["edges"], (name, r)
    assert batches["knn_edges_2p24"]["edges"] == 1 << 24 and batches["long_edges_2p20"]["edges"] == 1 << 20
    assert batches["long_edges_2p20"]["mean_probes_per_edge"] > 100  # up to 3859 px: marched by the whole warp


def test_c2_cdrt_600_scripted_ticks(ctx, oracle):
    """CDRT2/test/testD.cpp:197-241 on virtual time: 600 ticks (30 s at 20 Hz) with the four moving obstacles and the agent
    following its path; the engine's state is compared with the CPU port's tick after tick (paths) and at the end (all state)."""
    import cdrt_bench

    out = cdrt_bench.single(ctx, oracle, ticks=600)
    assert out["gpu_status"] == 0 and out["state_matches_cpu"], out
    if "state_matches_reference" in out:
        assert out["state_matches_reference"], out


def test_c3_rtfmt_1e5_samples_to_exhaustion(ctx, oracle):
    import config_runs as cr

    out = cr.c3_rtfmt(ctx, oracle)
    assert out["gpu_status"] == 0 and out["nodes"] > 90_000
    assert out["n_unv_left"] == 0 or out["z"] < 0, out  # the expansion ran to its end
    assert out["state_matches_cpu"], out


def test_c5_sweep_8192_instances_tree_bytes(ctx, oracle):
    """8 maps x 32 pairs x 32 seeds, 2000 samples each (bench.measure_sweep / measure_replanning): every 128th instance's tree is
    compared byte for byte with the CPU oracle's."""
    import torch

    import bench

    sweep = bench.measure_sweep(ctx, 0, 1, torch, None, torch.device("cuda", 0))
    o = sweep["out"]
    assert o["instances"] == 8192 and o["gpu_status_max"] == 0 and o["samples_consumed_min"] == 2000, o
    rep = bench.measure_replanning(ctx, sweep)["replan"]
    assert rep["trees_compared"] == 64 and rep["trees_match_gpu"], rep

"""Planner-engine parity: the GPU RT-RRT* engine (one warp = one planner instance) must reproduce
the oracle's sequential run bit for bit — node set, parents, costs, collision flags, queue sizes,
clock reads, path — under the same virtual clock and the same injected samples."""
import numpy as np
import pytest

from helpers import dyn_plane

pytestmark = pytest.mark.gpu


def _run_pair(ctx, oracle, name, agent, goal, dt, ticks, seed, n_inst=1, moves=True, device_sampler=False):
    import rtpp_b200

    S = oracle.load_map(name)
    H, W = S.shape
    max_samples = 16384 if device_sampler else 4096
    informed = fallbacks = 0
    eng = rtpp_b200.Engine(ctx, n_inst, 8192, 5, W, H, max_samples)
    eng.set_map(0, 1, S)
    orcs = []
    for i in range(n_inst):
        o = oracle.OrcRTRRT(S)
        o.set_clock(dt)
        assert o.set_agent(*agent) and o.set_goal(*goal)
        orcs.append(o)
    eng.init_instances(np.zeros(n_inst, np.int32), np.tile(np.array(agent, np.float32), (n_inst, 1)),
                       np.tile(np.array(goal, np.float32), (n_inst, 1)))
    streams = [oracle.rand31(seed + i, 3_000_000) for i in range(n_inst)]
    pos = [0] * n_inst
    rng = np.random.default_rng(seed)
    for tick in range(ticks):
        if tick in (2, 5):
            D = dyn_plane(S.shape, rng, n_blobs=6)
            eng.set_map(0, 2, D)
            for o in orcs:
                o.set_dobs(D)
        samples = np.zeros((n_inst, max_samples, 2), np.float32)
        ns = np.zeros(n_inst, np.int32)
        offs_all = []
        for i, o in enumerate(orcs):
            o.update_tree()
            pts, offs = o.presample(streams[i][pos[i]:], max_samples)
            samples[i, : len(pts)] = pts
            ns[i] = len(pts)
            offs_all.append(offs)
        if device_sampler:
            # the engine draws the samples itself from the raw rand() stream (SamplePoint on the device, glibc's sinf / cosf /
            # atan2f bits); the oracle's pre-drawn points above are what it must come up with
            chunk = np.zeros((n_inst, max_samples), np.int32)
            for i in range(n_inst):
                part = streams[i][pos[i]: pos[i] + max_samples]
                chunk[i, : len(part)] = part
            eng.set_rand(chunk, np.full(n_inst, max_samples, np.int32))
        else:
            eng.set_inputs(samples=samples, n_samples=ns)
        eng.tick(update_tree=True, run_ms=50, dt_ns=dt, get_path=True)
        st = eng.stats()
        for i, o in enumerate(orcs):
            used = o.expand_rewiring(50, samples[i, : ns[i]])
            if device_sampler:
                assert st[i][14] == int(offs_all[i][used]), f"tick {tick} inst {i}: rand() draws consumed"
            pos[i] += int(offs_all[i][used])
            opath, ocol, ocost, ocb = o.get_path()
            informed += int(ocb >= 0)  # a path was found: the next tick samples the informed ellipse
            a = o.dump()
            b = eng.read_tree(i)
            d = dict(zip(eng.STAT, st[i].tolist()))
            assert d["status"] == 0
            assert d["samples_used"] == used
            assert d["n_nodes"] == o.num_nodes()
            for x, y in zip(a, b):
                assert np.array_equal(x.view(np.uint8), y.view(np.uint8)), f"tick {tick} inst {i}"
            assert (d["qr_size"], d["qs_size"]) == o.queue_sizes()
            assert d["clock_reads"] == o.L.orc_rtrrt_clock_reads(o.h) - (0)  # GetPath reads no clock
            assert np.array_equal(b[0][d["root"]], o.root())
            # path: the goal-entry branches and the depth-5 fallback (RTRRT.cpp:401-437: chain of the reachable node nearest
            # the goal, reported with collision = 1 and cost 3.4e38), all on the device
            gcost = np.array([d["path_cost_bits"]], np.uint32).view(np.float32)[0]
            ids = eng.read_path(i)
            gp = np.array([goal if j == -2 else b[0][j] for j in ids], np.float32).reshape(-1, 2)
            assert d["path_col"] == int(bool(ocol))
            assert np.array_equal(gp, np.asarray(opath, np.float32).reshape(-1, 2)), f"tick {tick} inst {i}: path (collision {ocol})"
            assert np.array_equal(np.float32(gcost).view(np.uint32), np.float32(ocost).view(np.uint32))
            fallbacks += int(bool(ocol) and len(ids) > 0)
        if moves and tick in (3, 6):
            # the agent walks to some tree node: next UpdateTree re-roots there
            new_agents = np.zeros((n_inst, 2), np.float32)
            for i, o in enumerate(orcs):
                xy = o.dump()[0]
                t = xy[(len(xy) * 2) // 3]
                o.set_agent(float(t[0]), float(t[1]))
                new_agents[i] = t
            eng.set_inputs(agent_xy=new_agents)
    sizes = [o.num_nodes() for o in orcs]
    eng.close()
    _run_pair.fallbacks = fallbacks
    if device_sampler:
        return sizes, informed
    return sizes


def test_rtrrt_engine_single_instance_mess6(ctx, oracle):
    sizes = _run_pair(ctx, oracle, "MESS6", (720.0, 720.0), (70.0, 70.0), dt=5000, ticks=8, seed=11)
    assert sizes[0] > 200


def test_rtrrt_engine_game0_reaches_goal(ctx, oracle):
    sizes = _run_pair(ctx, oracle, "GAME0", (350.0, 2350.0), (1400.0, 2050.0), dt=2000, ticks=8, seed=3)
    assert sizes[0] > 300
    # the goal is out of reach in these 8 ticks: every GetPath took the depth-5 fallback (device: csrc/path_fallback.cuh)
    assert _run_pair.fallbacks >= 4


def test_rtrrt_engine_many_instances(ctx, oracle):
    sizes = _run_pair(ctx, oracle, "Trap", (660.0, 660.0), (1160.0, 1160.0), dt=20000, ticks=5, seed=100, n_inst=12)
    assert len(set(sizes)) > 1  # different seeds -> different trees


def test_rtrrt_engine_device_sampler_uniform_and_informed(ctx, oracle):
    """rtpp_engine_set_rand: the engine draws its samples on the device from the raw rand() stream (SamplePoint, RTRRT.cpp:234-262).
    On FLOOR5 (the scenario of test_oracle_ref.py::test_rtrrt_live_with_informed_sampling) a path to the goal is found after a
    few ticks, which switches 90 % of the samples to the informed ellipse (glibc's atan2f / cosf / sinf on the device): trees,
    paths, clock reads AND the number of rand() draws consumed must equal the oracle's."""
    sizes, informed = _run_pair(ctx, oracle, "FLOOR5", (1000.0, 1200.0), (1750.0, 900.0), dt=2000, ticks=24, seed=21, moves=False, device_sampler=True)
    assert sizes[0] > 300 and informed >= 2
    sizes, informed = _run_pair(ctx, oracle, "Trap", (660.0, 660.0), (1160.0, 1160.0), dt=20000, ticks=6, seed=100, n_inst=12, device_sampler=True)
    assert len(set(sizes)) > 1

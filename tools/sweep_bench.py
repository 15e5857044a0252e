"""Scenario sweep (BASELINE config 5 flavour): many independent RT-RRT* instances on the GPU engine
vs the CPU oracle on host threads.  Development aid + helper for bench.py extras."""
import os, sys, time, json
from concurrent.futures import ThreadPoolExecutor
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# start / goals from the reference's scripted scenarios (RTRRT/test/testD.cpp:36-47)
SCEN = {
    "FLOOR5": ((1000, 1200), [(400, 300), (1920, 600), (1750, 900), (1460, 2100)]),
    "GAME0": ((350, 2350), [(1400, 2050), (888, 800), (2500, 1500), (2000, 400)]),
    "MAZE0": ((1500, 1850), [(200, 200), (3100, 200), (3100, 3100), (200, 3100)]),
    "MESS6": ((720, 720), [(70, 70), (1240, 70), (1240, 1240), (70, 1240)]),
    "FLOOR2": ((1100, 2200), [(440, 110), (790, 1440), (1040, 220), (1650, 350)]),
    "GAME6": ((550, 2048), [(520, 980), (245, 320), (2130, 1500), (1980, 460)]),
    "MAZE5": ((1600, 1780), [(105, 160), (3100, 90), (150, 3010), (3100, 3000)]),
    "Trap": ((660, 660), [(1160, 1160), (130, 1160), (130, 130), (1160, 130)]),
}


def make_instances(n_inst, maps):
    inst = []
    names = list(maps)
    for i in range(n_inst):
        name = names[i % len(names)]
        start, goals = SCEN[name]
        inst.append((names.index(name), name, start, goals[(i // len(names)) % 4], 1000 + i))
    return inst


def c5_pairs(name, S):
    """The 32 (start, goal) pairs of a map for BASELINE configs[4]: 16 from the reference's scripted table —
    ordered pairs of its start and four goals (RTRRT/test/testD.cpp:36-47) — and 16 seeded random pairs of free pixels."""
    start, goals = SCEN[name]
    pts = [start] + list(goals)
    table = [(a, b) for a in pts for b in pts if a != b][:16]
    rng = np.random.Generator(np.random.PCG64(sum(map(ord, name))))
    free = np.argwhere(S == 0)
    rnd = []
    while len(rnd) < 16:
        (ay, ax), (by, bx) = free[rng.integers(len(free))], free[rng.integers(len(free))]
        if min(ax, ay, bx, by) > 2 and (ax, ay) != (bx, by):
            rnd.append(((int(ax), int(ay)), (int(bx), int(by))))
    return table + rnd


def make_instances_c5(maps, n_seeds=32):
    """8 maps x 32 (start, goal) pairs x n_seeds sample streams; instance id = ((seed * 32) + pair) * 8 + map, so that any
    contiguous or strided subset mixes maps and pairs.  Entries as make_instances: (map slot, map, start, goal, rand seed)."""
    names = list(maps)
    pairs = {name: c5_pairs(name, maps[name]) for name in names}
    inst = []
    for seed in range(n_seeds):
        for pr in range(32):
            for slot, name in enumerate(names):
                a, b = pairs[name][pr]
                inst.append((slot, name, a, b, 5000 + (seed * 32 + pr) * len(names) + slot))
    return inst


def uniform_samples(orc, seed, n, W, H):
    """n points exactly as the reference's uniform branch draws them (p0 discarded, then x, y)."""
    r = orc.rand31(seed, 3 * n).astype(np.float32) / np.float32(2147483648.0)
    return np.stack([r[1::3] * np.float32(W), r[2::3] * np.float32(H)], 1).astype(np.float32)


def run_gpu(ctx, orc, inst, maps, n_samples, ticks, dt, run_ms=50, cap=4096, until_exhausted=False, keep=False):
    """ticks replanning ticks of every instance; until_exhausted: tick until every instance has consumed its n_samples samples
    (at most `ticks`).  keep: returns the live engine as a 4th value (the caller reads trees and closes it)."""
    import rtpp_b200
    W = max(m.shape[1] for m in maps.values()); H = max(m.shape[0] for m in maps.values())
    n = len(inst)
    eng = rtpp_b200.Engine(ctx, n, cap, 5, W, H, n_samples)
    for slot, (name, S) in enumerate(maps.items()):
        eng.set_map(slot, 1, S)
    eng.init_instances(np.array([i[0] for i in inst], np.int32), np.array([i[2] for i in inst], np.float32),
                       np.array([i[3] for i in inst], np.float32))
    samples = np.stack([uniform_samples(orc, i[4], n_samples, maps[i[1]].shape[1], maps[i[1]].shape[0]) for i in inst])
    eng.set_inputs(samples=samples, n_samples=np.full(n, n_samples, np.int32))
    ctx.sync()
    lat = []
    for t in range(ticks):
        t0 = time.perf_counter()
        eng.tick(update_tree=True, run_ms=run_ms, dt_ns=dt, get_path=True)
        ctx.sync()
        lat.append(time.perf_counter() - t0)
        if until_exhausted and int(eng.stats()[:, 2].min()) >= n_samples:
            break
    st = eng.stats()
    if keep:
        return lat, st, samples, eng
    eng.close()
    return lat, st, samples


def run_cpu_instance(orc, S, start, goal, samples, ticks, dt, run_ms=50, want_tree=False):
    """The CPU oracle port of the same loop; `ticks` ticks exactly (the GPU run's count).  want_tree: also returns o.dump()."""
    o = orc.OrcRTRRT(S)
    o.set_clock(dt)
    o.set_agent(float(start[0]), float(start[1])); o.set_goal(float(goal[0]), float(goal[1]))
    pos = 0
    lat = []
    for t in range(ticks):
        t0 = time.perf_counter()
        o.update_tree()
        pos += o.expand_rewiring(run_ms, samples[pos:])
        o.get_path()
        lat.append(time.perf_counter() - t0)
    c = o.counters(); c["n_nodes"] = o.num_nodes(); c["samples_used"] = pos
    tree = o.dump() if want_tree else None
    o.close()
    return (lat, c, tree) if want_tree else (lat, c)


def main():
    import rtpp_b200
    from oracle import loader as orc
    n_inst = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    dt = int(sys.argv[3]) if len(sys.argv) > 3 else 5000
    n_samples = 2000
    maps = {name: orc.load_map(name) for name in SCEN}
    inst = make_instances(n_inst, maps)
    ctx = rtpp_b200.Ctx(0)
    lat, st, samples = run_gpu(ctx, orc, inst, maps, n_samples, ticks, dt)
    tot = sum(lat)
    print(json.dumps({"gpu": {"instances": n_inst, "ticks": ticks, "dt_ns": dt, "tick_ms": [round(x * 1e3, 2) for x in lat],
                              "iters_per_s": float(st[:, 4].sum() / tot), "pops_per_s": float((st[:, 5] + st[:, 6]).sum() / tot),
                              "nodes_mean": float(st[:, 0].mean()), "status_bad": int((st[:, 9] != 0).sum())}}))
    # CPU: the oracle port on all host threads over a sample of the instances
    threads = os.cpu_count() or 1
    m = min(n_inst, 4 * threads)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        res = list(ex.map(lambda k: run_cpu_instance(orc, maps[inst[k][1]], inst[k][2], inst[k][3], samples[k], ticks, dt), range(m)))
    wall = time.perf_counter() - t0
    iters = sum(r[1]["iters"] for r in res); pops = sum(r[1]["pops_r"] + r[1]["pops_s"] for r in res)
    same = all(res[k][1]["n_nodes"] == int(st[k, 0]) and res[k][1]["samples_used"] == int(st[k, 2]) for k in range(m))
    p50 = float(np.median([x for r in res for x in r[0]]))
    print(json.dumps({"cpu_port": {"instances": m, "threads": threads, "wall_s": round(wall, 3), "iters_per_s": iters / wall,
                                   "pops_per_s": pops / wall, "single_thread_tick_p50_ms": round(p50 * 1e3, 2), "matches_gpu": same}}))
    # single-instance latency on the GPU
    lat1, st1, _ = run_gpu(ctx, orc, inst[:1], maps, n_samples, max(ticks, 8), dt)
    print(json.dumps({"gpu_single": {"tick_ms": [round(x * 1e3, 2) for x in lat1], "p50_ms": round(float(np.median(lat1)) * 1e3, 2),
                                     "pops": int(st1[0, 5] + st1[0, 6]), "nodes": int(st1[0, 0])}}))


if __name__ == "__main__":
    main()

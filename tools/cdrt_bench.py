#!/usr/bin/env python
"""CDRT-RRT* (BASELINE configs[1]): replanning tick latency and multi-instance throughput of the GPU engine against
the CPU oracle port on the same scripted scenario (virtual clock, injected rand).  Prints one JSON object."""
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cdrt_scenarios as sc  # noqa: E402


def single(ctx, orc, name="GAME0", goal=2, seed=7, dt=2000, ticks=60):
    import rtpp_b200

    fx = orc.load_cdrt_fixture(name)
    r = orc.rand31(seed, 2_000_000)
    o = orc.CdrtCpu(fx, "oracle")
    g = rtpp_b200.CdrtPlanner(ctx, fx)
    try:  # the reference's own BImap (CDRT.cpp compiled in place; cv::Mat and CGAL k-NN are stand-ins), when it was built
        ref = orc.CdrtCpu(fx, "ref")
    except (FileNotFoundError, OSError):
        ref = None
    everyone = [p for p in (o, g, ref) if p is not None]
    for p in everyone:
        sc.setup(p, name, goal, r, dt_ns=dt)
    tc, tg, tr = [], [], []
    same = same_ref = True
    for k in range(ticks):
        for p, acc in ((o, tc), (g, tg), (ref, tr)):
            if p is None:
                continue
            t0 = time.perf_counter()
            p.task()
            acc.append(time.perf_counter() - t0)
        pa, pb = o.get_path(), g.get_path()
        same &= bool(np.array_equal(pa[0], pb[0]) and pa[2] == pb[2])
        if ref is not None:
            pr = ref.get_path()
            same_ref &= bool(np.array_equal(pr[0], pb[0]) and pr[2] == pb[2])
        t = (k + 1) * sc.TICK_S
        for i, wp in enumerate(sc.OBS_PATHS[name]):
            q = sc.along(wp, sc.OBS_SPEED * t)
            for p in everyone:
                p.move_obstacle(i, float(q[0]), float(q[1]), float(sc.OBSTACLES[name][i][3]))
        if len(pa[0]) >= 2:
            a = sc.along(pa[0], sc.SPEED * sc.TICK_S)
            for p in everyone:
                p.set_agent(float(a[0]), float(a[1]))
    da, db = o.dump(), g.dump()
    same &= all(np.array_equal(x.view(np.uint8), y.view(np.uint8)) for x, y in zip(da, db))
    st = g.state()
    out = {"map": name, "ticks": ticks, "virtual_dt_ns": dt, "pops_per_tick": st["pops"] / ticks, "pushes_per_tick": st["pushes"] / ticks,
           "nodes": len(da[0]), "gpu_tick_ms_p50": float(np.median(tg) * 1e3), "cpu_tick_ms_p50": float(np.median(tc) * 1e3),
           "cpu_kind": "port (oracle/cdrt_oracle.hpp), 1 thread", "state_matches_cpu": bool(same), "gpu_status": st["status"]}
    # most of the 600 scripted ticks are light (no obstacle near the tree): p50 alone hides the heavy replanning ticks
    def spread(key, ts):
        a = np.asarray(ts) * 1e3
        out[key + "_mean"], out[key + "_p90"], out[key + "_max"] = float(a.mean()), float(np.percentile(a, 90)), float(a.max())

    spread("gpu_tick_ms", tg)
    spread("cpu_tick_ms", tc)
    if ref is not None:
        spread("reference_tick_ms", tr)
        dr = ref.dump()
        same_ref &= all(np.array_equal(x.view(np.uint8), y.view(np.uint8)) for x, y in zip(dr, db))
        out["reference_tick_ms_p50"] = float(np.median(tr) * 1e3)
        out["reference_kind"] = "BImap::RTRT_task of the reference's CDRT.cpp compiled in place (oracle/_ref), 1 thread"
        out["state_matches_reference"] = bool(same_ref)
        ref.close()
    o.close()
    g.close()
    return out


def sweep(ctx, orc, name="GAME0", n_inst=128, ticks=20, dt=20000, cpu_inst=16, threads=16):
    import rtpp_b200

    fx = orc.load_cdrt_fixture(name)
    start, goals = sc.TASKS[name]
    eng = rtpp_b200.CdrtEngine(ctx, n_inst, fx["graphs"][0], fx["amap"], fx["imap"], fx["resolution"], queue_capacity=1 << 19,
                               rand_capacity=400_000)

    def setup_cpu(i):
        o = orc.CdrtCpu(fx, "oracle")
        o.set_params()
        o.set_clock(dt)
        o.set_rand(orc.rand31(1000 + i, 400_000))
        for k, (typ, x, y, ang, rad, outline) in enumerate(sc.OBSTACLES[name]):
            o.add_obstacle(k, typ, float(x), float(y), float(ang), float(rad), np.array(outline, np.float64), 2.0)
        o.set_agent(float(start[0]), float(start[1]))
        o.set_goal(*[float(v) for v in goals[i % 4]])
        o.start()
        return o

    for i in range(n_inst):
        for k, (typ, x, y, ang, rad, outline) in enumerate(sc.OBSTACLES[name]):
            eng.add_obstacle(i, k, typ, float(x), float(y), float(ang), float(rad), np.array(outline, np.float64), 2.0)
        eng.set_agent(i, float(start[0]), float(start[1]))
        eng.set_goal(i, *[float(v) for v in goals[i % 4]])
        eng.start(i, orc.rand31(1000 + i, 400_000))
    cpus = [setup_cpu(i) for i in range(cpu_inst)]

    def obstacle_poses(k):
        t = (k + 1) * sc.TICK_S
        return [sc.along(wp, sc.OBS_SPEED * t) for wp in sc.OBS_PATHS[name]]

    tg = []
    for k in range(ticks):
        t0 = time.perf_counter()
        eng.tick(dt, get_path=True)
        tg.append(time.perf_counter() - t0)
        for j, q in enumerate(obstacle_poses(k)):
            for i in range(n_inst):
                eng.move_obstacle(i, j, float(q[0]), float(q[1]), float(sc.OBSTACLES[name][j][3]))
    gs = eng.stats()

    def run_cpu(o):
        for k in range(ticks):
            o.task()
            o.get_path()
            for j, q in enumerate(obstacle_poses(k)):
                o.move_obstacle(j, float(q[0]), float(q[1]), float(sc.OBSTACLES[name][j][3]))
        return o.counters()

    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        cs = list(ex.map(run_cpu, cpus))
    cpu_s = time.perf_counter() - t0
    same = True
    for i, o in enumerate(cpus):
        pl = rtpp_b200.CdrtPlanner(ctx, fx, inst=i, engine=eng)
        same &= all(np.array_equal(x.view(np.uint8), y.view(np.uint8)) for x, y in zip(o.dump(), pl.dump()))
        same &= (cs[i]["pops"], cs[i]["pushes"]) == (int(gs[i, 12]), int(gs[i, 15]))
        o.close()
    work_g = int(gs[:, 12].sum() + gs[:, 14].sum())
    work_c = sum(c["pops"] + c["sampled"] for c in cs)
    out = {"map": name, "instances": n_inst, "ticks": ticks, "virtual_dt_ns": dt, "gpu_status_max": int(gs[:, 11].max()),
           "gpu_sweep_tick_ms_p50": float(np.median(tg) * 1e3), "gpu_instance_ticks_per_s": n_inst * ticks / sum(tg),
           "gpu_iters_per_s": work_g / sum(tg), "cpu_instances": cpu_inst, "cpu_threads": threads,
           "cpu_instance_ticks_per_s": cpu_inst * ticks / cpu_s, "cpu_iters_per_s": work_c / cpu_s,
           "iteration": "one eQr pop or one accepted sample", "trees_match_cpu": bool(same)}
    eng.close()
    return out


if __name__ == "__main__":
    import rtpp_b200
    from oracle import loader as orc

    with rtpp_b200.Ctx(0) as ctx:
        n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
        res = {"single": single(ctx, orc), "sweep": sweep(ctx, orc, n_inst=n)}
    print(json.dumps(res))

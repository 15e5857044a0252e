#!/usr/bin/env python
"""Where one CDRT-RRT* planner's tick goes on the GPU engine (development aid): per-phase SM cycles of the scripted
GAME0 scenario of tools/cdrt_bench.py, GPU only."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cdrt_scenarios as sc  # noqa: E402
import rtpp_b200  # noqa: E402
from oracle import loader as orc  # noqa: E402

name, goal, seed, dt, ticks = "GAME0", 2, 7, 2000, int(sys.argv[1]) if len(sys.argv) > 1 else 40
ctx = rtpp_b200.Ctx(0)
fx = orc.load_cdrt_fixture(name)
g = rtpp_b200.CdrtPlanner(ctx, fx)
sc.setup(g, name, goal, orc.rand31(seed, 2_000_000), dt_ns=dt)
prev = g.state()
for k in range(ticks):
    t0 = time.perf_counter()
    g.task()
    ms = (time.perf_counter() - t0) * 1e3
    st = g.state()
    fine = g.eng.read_cycles(g.inst)
    pa = g.get_path()
    t = (k + 1) * sc.TICK_S
    for i, wp in enumerate(sc.OBS_PATHS[name]):
        q = sc.along(wp, sc.OBS_SPEED * t)
        g.move_obstacle(i, float(q[0]), float(q[1]), float(sc.OBSTACLES[name][i][3]))
    if len(pa[0]) >= 2:
        a = sc.along(pa[0], sc.SPEED * sc.TICK_S)
        g.set_agent(float(a[0]), float(a[1]))
    if k % 5 == 4 or k == ticks - 1:
        d = {x: st[x] - prev[x] for x in ("pops", "pushes", "edge_checks", "rewired", "sampled")}
        cyc = {x: st[x] for x in ("cyc_env_tree", "cyc_sets", "cyc_sampling", "cyc_queue")}
        print(f"tick {k}: {ms:.2f} ms queue={st['queue']} nodes={st['n_ids']} {d}")
        print("   kcycles:", {x: round(v / 1e3) for x, v in cyc.items()}, {x: (round(v / 1e3) if x.startswith('cyc') else v) for x, v in fine.items()})
    prev = st

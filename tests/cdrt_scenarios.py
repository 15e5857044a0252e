"""CDRT-RRT* scripted scenarios: the tables of the reference demo CDRT2/test/testD.cpp:50-97 (start, goals,
dynamic obstacles and their way-point paths) and a deterministic re-statement of its main loop
(testD.cpp:117-228) on VIRTUAL time: one tick = 50 ms (cv::waitKey(50)), obstacles move at 100 px/s along
their polylines, the agent follows the planner's own path at 400 px/s.  The loop below is pure Python on
binary64 and drives any planner object with the CdrtCpu call surface (reference, oracle, GPU engine)."""
import numpy as np

SPEED, OBS_SPEED, TICK_S = 400.0, 100.0, 0.05

# name: (start, goals)                                                         testD.cpp:52-55
TASKS = {
    "FLOOR5": ((1000, 1200), [(400, 300), (1920, 600), (1750, 900), (1460, 2100)]),
    "GAME0": ((350, 2350), [(1400, 2050), (888, 800), (2500, 1500), (2000, 400)]),
    "MAZE0": ((1500, 1850), [(200, 200), (3100, 200), (3100, 3100), (200, 3100)]),
    "MESS6": ((720, 720), [(70, 70), (1240, 70), (1240, 1240), (70, 1240)]),
}
# (type, x, y, angle, radius, outline)                                          testD.cpp:57-76
OBSTACLES = {
    "MESS6": [(1, 284, 284, 0, 0, [(-65, 65), (65, 65), (65, -65), (-65, -65)]),
              (1, 1034, 198, 0, 0, [(-54, -28), (-28, -28), (-28, -54), (28, -54), (28, -28), (54, -28), (54, 28), (28, 28), (28, 54),
                                    (-28, 54), (-28, 28), (-54, 28)]),
              (1, 1044, 889, 0, 0, [(0, -81), (-81, 81), (81, 81)]),
              (0, 350, 955, 0, 82, [(0, 0)])],
    "MAZE0": [(0, 485, 1040, 0, 70, [(0, 0)]), (0, 2650, 520, 0, 70, [(0, 0)]), (0, 1930, 2750, 0, 70, [(0, 0)]),
              (0, 850, 2780, 0, 70, [(0, 0)])],
    "GAME0": [(0, 370, 388, 0, 57, [(0, 0)]), (0, 2370, 630, 0, 114, [(0, 0)]),
              (1, 1364, 1364, float(np.float32(0.78539816)), 0, [(0, -57), (57, 0), (57, 57), (-57, 57), (-57, 0)]),
              (1, 1340, 2520, 0, 0, [(-57, -57), (57, -57), (57, 57), (-57, 57)])],
    "FLOOR5": [(0, 535, 435, 0, 49, [(0, 0)]), (1, 425, 1750, 0, 0, [(-49, -49), (49, -49), (49, 49), (-49, 49)]),
               (1, 1246, 2022, 0, 0, [(0, -49), (49, 49), (-49, 49)]),
               (1, 1378, 888, 0, 0, [(-52, 0), (-31, -49), (31, -49), (52, 0), (31, 49), (-31, 49)])],
}
# way-points of each obstacle                                                   testD.cpp:78-97
OBS_PATHS = {
    "MESS6": [[(284, 284), (880, 230), (777, 530)], [(1034, 198), (1200, 920), (747, 980)], [(1044, 889), (1044, 425)],
              [(350, 955), (60, 630), (260, 447)]],
    "MAZE0": [[(485, 1040), (465, 1426), (170, 1420), (180, 870)], [(2650, 520), (2420, 525), (2420, 1130), (1600, 1130)],
              [(1930, 2750), (2100, 2760), (2100, 2150), (1800, 2180), (1800, 2500)], [(850, 2780), (850, 2400), (100, 2400)]],
    "GAME0": [[(370, 388), (225, 630), (2230, 1460)], [(2370, 630), (2340, 240), (2000, 200)], [(1364, 1364), (1900, 850)],
              [(1340, 2520), (1490, 2500), (1480, 2350), (1760, 2260)]],
    "FLOOR5": [[(535, 435), (700, 490), (700, 1130)], [(425, 1750), (240, 1800), (400, 1650), (580, 1850)],
               [(1246, 2022), (1300, 1970), (1300, 1750), (1900, 1740)], [(1378, 888), (1300, 930), (1025, 900)]],
}


def along(path, dlen):
    """Point `dlen` along a polyline from its first vertex (testD.cpp:176-192): stops at the last vertex."""
    path = np.asarray(path, np.float64)
    p_old = path[0]
    for p in path:
        ln = float(np.sqrt((p[0] - p_old[0]) * (p[0] - p_old[0]) + (p[1] - p_old[1]) * (p[1] - p_old[1])))
        if ln < dlen:
            dlen -= ln
        else:
            t = dlen / ln if ln > 0 else 0.0
            return p * t + p_old * (1 - t)
        p_old = p
    return path[-1].copy()


def setup(pl, name, goal_idx, seed_stream, dt_ns=20000, params=None):
    """BImap set-up of testD.cpp:109-122 on planner `pl`."""
    start, goals = TASKS[name]
    pl.set_params(**(params or {}))
    pl.set_clock(dt_ns)
    pl.set_rand(seed_stream)
    for i, (typ, x, y, ang, rad, outline) in enumerate(OBSTACLES[name]):
        pl.add_obstacle(i, typ, float(x), float(y), float(ang), float(rad), np.array(outline, np.float64), 2.0)
    pl.set_agent(float(start[0]), float(start[1]))
    pl.set_goal(float(goals[goal_idx][0]), float(goals[goal_idx][1]))
    pl.start()


def tick(pl, name, k):
    """One iteration of the demo loop at virtual time (k+1)*TICK_S: task, GetPath, move obstacles, move the agent."""
    pl.task()
    path, col, cost = pl.get_path()
    t = (k + 1) * TICK_S
    for i, wp in enumerate(OBS_PATHS[name]):
        p = along(wp, OBS_SPEED * t)
        pl.move_obstacle(i, float(p[0]), float(p[1]), float(OBSTACLES[name][i][3]))
    if len(path) >= 2:
        a = along(path, SPEED * TICK_S)
        pl.set_agent(float(a[0]), float(a[1]))
    return path, col, cost


# golden runs (tools/make_cdrt_golden.py -> tests/golden/cdrt_<name>.npz, from the compiled reference)
GOLDEN_CASES = [
    dict(name="mess6_g0", map="MESS6", goal=0, seed=7, n_rand=400_000, dt=20000, ticks=40, dump_ticks=[0, 9, 24, 39],
         params=dict(r_th=10.0, p_th=0.85, nearK=20, RewireN=100)),
    dict(name="floor5_g3", map="FLOOR5", goal=3, seed=11, n_rand=400_000, dt=10000, ticks=30, dump_ticks=[0, 14, 29],
         params=dict(r_th=10.0, p_th=0.85, nearK=20, RewireN=100)),
]


def dump_hash(d):
    import hashlib

    h = hashlib.sha256()
    for a in d:
        h.update(np.ascontiguousarray(a).tobytes())
    return np.frombuffer(h.digest(), np.uint8).copy()

import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rtpp_b200
from oracle import loader as oracle
from helpers import dyn_plane

name, agent, goal, dt, ticks, seed = "MESS6", (720.0, 720.0), (70.0, 70.0), 5000, 8, 11
ctx = rtpp_b200.Ctx(0)
S = oracle.load_map(name); H, W = S.shape
max_samples = 4096
eng = rtpp_b200.Engine(ctx, 1, 8192, 5, W, H, max_samples)
eng.set_map(0, 1, S)
o = oracle.OrcRTRRT(S); o.set_clock(dt); o.set_agent(*agent); o.set_goal(*goal)
eng.init_instances(np.zeros(1, np.int32), np.array([agent], np.float32), np.array([goal], np.float32))
r = oracle.rand31(seed, 3_000_000); pos = 0
rng = np.random.default_rng(seed)
for tick in range(ticks):
    if tick in (2, 5):
        D = dyn_plane(S.shape, rng, n_blobs=6); eng.set_map(0, 2, D); o.set_dobs(D)
    o.update_tree()
    pts, offs = o.presample(r[pos:], max_samples)
    samples = np.zeros((1, max_samples, 2), np.float32); samples[0, :len(pts)] = pts
    eng.set_inputs(samples=samples, n_samples=np.array([len(pts)], np.int32))
    eng.tick(update_tree=True, run_ms=50, dt_ns=dt, get_path=True)
    used = o.expand_rewiring(50, pts); pos += int(offs[used])
    op = o.get_path()
    d = eng.stats_dict(0)
    a = o.dump(); b = eng.read_tree(0)
    n = min(len(a[0]), len(b[0]))
    same_xy = np.array_equal(a[0][:n], b[0][:n])
    first_diff = -1
    if not same_xy:
        first_diff = int(np.flatnonzero((a[0][:n] != b[0][:n]).any(1))[0])
    print("  roots", o.root(), b[0][d["root"]], "parent-mismatch idx", np.flatnonzero(a[1][:n] != b[1][:n])[:10], "cost-mismatch idx", np.flatnonzero(a[2][:n].view(np.uint32) != b[2][:n].view(np.uint32))[:10])
    print(tick, "orc n", o.num_nodes(), "eng", d, "used", used, "orc queues", o.queue_sizes(), "clk", o.L.orc_rtrrt_clock_reads(o.h), "first xy diff", first_diff,
          "par same", np.array_equal(a[1][:n], b[1][:n]), "cost same", np.array_equal(a[2][:n].view(np.uint32), b[2][:n].view(np.uint32)), "orc path", op[1:], len(op[0]))
    if first_diff >= 0:
        print(" orc node", a[0][first_diff], "eng node", b[0][first_diff]); break
    if o.num_nodes() != d["n_nodes"]: break
    if tick in (3, 6):
        xy = a[0]; t = xy[(len(xy) * 2) // 3]
        print("  move agent to node", (len(xy) * 2) // 3, t, "set_agent ok:", o.set_agent(float(t[0]), float(t[1])))
        eng.set_inputs(agent_xy=np.array([t], np.float32))

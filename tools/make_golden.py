#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref = the reference's planner
sources compiled in place, see oracle/Makefile).  Runs only where /root/reference exists; the
vectors are committed so that the oracle can be pinned anywhere (tests/test_oracle_golden.py).

Inputs are regenerated in the tests from the same seeded helpers (tests/helpers.py), so only the
reference's OUTPUTS are stored.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import dyn_plane, edge_mix, golden_cases  # noqa: E402
from oracle import loader as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def edges_case(name):
    S = O.load_map(name)
    H, W = S.shape
    D = dyn_plane(S.shape, np.random.default_rng(7))
    e = edge_mix(np.random.default_rng(123), 200_000, W, H)
    rt = O.RefRTRRT(S); rt.set_dobs(D)
    fm = O.RefRTFMT(S); fm.set_dobs(D)
    am = O.RefAMRRT(S); am.set_dobs(D)
    e1 = rt.crash(e)
    assert np.array_equal(e1, fm.crash(e))  # RTFMT.cpp:242-262 is the same routine
    e2 = fm.crash(e, ndobs=True)
    e3 = am.crash(e)
    e4b, e4p = am.crash_start(e[:20_000])
    dist = rt.dist(e[:20_000])
    np.savez_compressed(os.path.join(OUT, f"edges_{name}.npz"), e1=np.packbits(e1), e2=np.packbits(e2), e3=np.packbits(e3),
                        e4_blocked=np.packbits(e4b), e4_pstart=e4p, dist=dist)
    print("edges", name, e1.mean(), e2.mean(), e3.mean(), e4b.mean())


def rtrrt_case(case):
    S = O.load_map(case["map"])
    ref = O.RefRTRRT(S)
    r = O.rand31(case["seed"], case["n_rand"])
    ref.set_rand(r)
    ref.set_clock(case["dt"], 0)
    ref.set_agent(*case["agent"]); ref.set_goal(*case["goal"])
    rng = np.random.default_rng(case["seed"])
    dumps = []
    for tick in range(case["ticks"]):
        if tick in case["dobs_ticks"]:
            ref.set_dobs(dyn_plane(S.shape, rng, n_blobs=6))
        ref.update_tree()
        ref.expand_rewiring(50)
        path, col, cost, cb = ref.get_path()
        xy, par, c, cl = ref.dump()
        dumps.append(dict(xy=xy, par=par, cost=c, col=cl, path=path, path_col=col, path_cost=np.float32(cost), c_base=np.float32(cb),
                          rand_pos=ref.rand_pos(), queues=np.array(ref.queue_sizes())))
        if tick in case["move_ticks"]:
            t = xy[(len(xy) * 2) // 3]
            ref.set_agent(float(t[0]), float(t[1]))
    flat = {}
    for i, d in enumerate(dumps):
        for k, v in d.items():
            flat[f"t{i}_{k}"] = v
    np.savez_compressed(os.path.join(OUT, f"rtrrt_{case['name']}.npz"), **flat)
    print("rtrrt", case["name"], [len(d["xy"]) for d in dumps], [bool(d["path_col"]) for d in dumps])


def rtfmt_case(case):
    S = O.load_map(case["map"])
    ref = O.RefRTFMT(S)
    r = O.rand31(case["seed"], case["n_rand"])
    ref.set_agent(*case["agent"]); ref.set_goal(*case["goal"])
    used = ref.start(r, case["samples"])
    rng = np.random.default_rng(case["seed"])
    flat = {"rand_used": used}
    for rnd in range(case["rounds"]):
        if rnd in case["dobs_rounds"]:
            ref.set_dobs(dyn_plane(S.shape, rng, n_blobs=5))
        ref.update_tree()
        ref.expand(500, True)
        xy, par, c, cl, st = ref.dump()
        flat.update({f"r{rnd}_par": par, f"r{rnd}_cost": c, f"r{rnd}_col": cl, f"r{rnd}_state": st, f"r{rnd}_z": ref.z(),
                     f"r{rnd}_path": ref.get_path()})
        if rnd == 0:
            flat["xy"] = xy
        if rnd in case["move_rounds"]:
            p = ref.get_path()
            if len(p) > 2:
                ref.set_agent(float(p[1][0]), float(p[1][1]))
    np.savez_compressed(os.path.join(OUT, f"rtfmt_{case['name']}.npz"), **flat)
    print("rtfmt", case["name"], used, np.bincount(flat[f"r{case['rounds'] - 1}_state"] & 3, minlength=4))


def amrrt_case(case):
    S = O.load_map(case["map"])
    H, W = S.shape
    ref = O.RefAMRRT(S)
    cells, mat, ratio = case["diffusion"](W)
    ref.set_diffusion(cells, mat, ratio)
    ref.set_rand(O.rand31(case["seed"], case["n_rand"]))
    ref.set_clock(case["dt"], 0)
    ref.set_agent(*case["agent"]); ref.set_goal(*case["goal"])
    rng = np.random.default_rng(case["seed"])
    flat = {}
    for tick in range(case["ticks"]):
        if tick in case["dobs_ticks"]:
            ref.set_dobs(dyn_plane(S.shape, rng, n_blobs=4))
        ref.main_task()
        xy, par, c, cl = ref.dump()
        path, pcol, pcost = ref.get_path()
        flat.update({f"t{tick}_xy": xy, f"t{tick}_par": par, f"t{tick}_cost": c, f"t{tick}_col": cl, f"t{tick}_rand_pos": ref.rand_pos(),
                     f"t{tick}_path": path, f"t{tick}_path_cost": np.float32(pcost), f"t{tick}_path_col": pcol})
    np.savez_compressed(os.path.join(OUT, f"amrrt_{case['name']}.npz"), **flat)
    print("amrrt", case["name"], [len(flat[f"t{t}_xy"]) for t in range(case["ticks"])])


def main():
    O.build(ref=True)
    only = set(sys.argv[1:])  # e.g. `make_golden.py floor5 maze0 FLOOR5 MAZE0`: regenerate just these cases / edge maps
    for name in ("GAME0", "MESS6", "FLOOR5", "MAZE0"):
        if not only or name in only:
            edges_case(name)
    cases = golden_cases()
    if only:
        cases = {k: [c for c in v if c["name"] in only] for k, v in cases.items()}
    for c in cases["rtrrt"]:
        rtrrt_case(c)
    for c in cases["rtfmt"]:
        rtfmt_case(c)
    for c in cases["amrrt"]:
        amrrt_case(c)


if __name__ == "__main__":
    main()

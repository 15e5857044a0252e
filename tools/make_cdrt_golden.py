#!/usr/bin/env python
"""tests/golden/cdrt_<map>.npz from the REFERENCE ITSELF: oracle/_ref/libref_cdrt.so (CDRT.cpp compiled in
place) driven through the scripted demo loop of tests/cdrt_scenarios.py.  Build container only.
Stored per tick: node count, queue length, rand/clock positions, path + cost, SHA-256 of the node dump;
full node dumps at a few ticks.  Inputs (graph, planes) are the fixtures of tools/make_cdrt_fixtures.py."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cdrt_scenarios as sc  # noqa: E402
from oracle import loader as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
CASES = sc.GOLDEN_CASES


def dump_hash(d):
    h = hashlib.sha256()
    for a in d:
        h.update(np.ascontiguousarray(a).tobytes())
    return np.frombuffer(h.digest(), np.uint8).copy()


def main():
    for case in CASES:
        fx = O.load_cdrt_fixture(case["map"])
        ref = O.CdrtCpu(fx, "ref")
        sc.setup(ref, case["map"], case["goal"], O.rand31(case["seed"], case["n_rand"]), dt_ns=case["dt"], params=case["params"])
        out = {}
        sxy, scut = ref.samples()
        out["samples_xy"], out["samples_cut"] = sxy, scut
        for i, a in enumerate(ref.dump()):
            out[f"start_dump{i}"] = a
        rows = []
        for k in range(case["ticks"]):
            path, col, cost = sc.tick(ref, case["map"], k)
            d = ref.dump()
            st = ref.state()
            rows.append((len(d[0]), st["queue"], ref.rand_pos(), ref.clock_reads(), st["island"], st["n_trees"], st["samples_left"],
                         st["larger"], st["smaller"], int(col), len(path)))
            out[f"t{k}_hash"] = dump_hash(d)
            out[f"t{k}_path"] = path
            out[f"t{k}_cost"] = np.float64(cost)
            if k in case["dump_ticks"]:
                for i, a in enumerate(d):
                    out[f"t{k}_dump{i}"] = a
                isl, cnt = ref.tree_islands()
                out[f"t{k}_islands"], out[f"t{k}_island_pts"] = isl, cnt
        out["rows"] = np.array(rows, np.int64)
        np.savez_compressed(os.path.join(OUT, f"cdrt_{case['name']}.npz"), **out)
        print(case["name"], "ticks", case["ticks"], "final nodes", rows[-1][0], "queue", rows[-1][1], "rand", rows[-1][2],
              os.path.getsize(os.path.join(OUT, f"cdrt_{case['name']}.npz")), "bytes")
        ref.close()


if __name__ == "__main__":
    main()

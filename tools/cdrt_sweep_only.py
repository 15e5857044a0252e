"""CDRT-RRT* sweep leg alone (development aid): python tools/cdrt_sweep_only.py [instances] [ticks]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rtpp_b200, cdrt_bench
from oracle import loader as orc

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    with rtpp_b200.Ctx(0) as ctx:
        print(json.dumps(cdrt_bench.sweep(ctx, orc, n_inst=n, ticks=ticks)))

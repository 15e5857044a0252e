"""RT-FMT at 1e5 samples (BASELINE configs[2]) on the engine vs the CPU port — development aid (bench.measure_fmt alone)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import rtpp_b200

if __name__ == "__main__":
    with rtpp_b200.Ctx(0) as ctx:
        print(json.dumps(bench.measure_fmt(ctx)))

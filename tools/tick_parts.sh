#!/bin/bash
# times the windowed tick's parts on the C4 workload (profiling aid)
run() { echo "$1: $(env $1 python tools/tick_bench.py 20 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["tick_window_ms_flushed"], d["tick_window_ms_warm"], d["tree_equal"], d["rewire_equal"])')"; }
for v in "RTPP_TICK_PARTS=1" "RTPP_TICK_PARTS=2" "RTPP_TICK_PARTS=3" "RTPP_TICK_BULK=0 RTPP_TICK_PARTS=3"; do run "$v"; done

#!/bin/bash
# times the windowed tick's parts on the C4 workload (profiling aid)
run() { echo "$1: $(env $1 python tools/tick_bench.py 20 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["tick_window_ms_flushed"], d["tick_window_ms_warm"], d["tree_equal"], d["rewire_equal"])')"; }
for v in "$@"; do run "$v"; done

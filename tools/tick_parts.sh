#!/bin/bash
# times the windowed tick's parts on the C4 workload (profiling aid): tree edges only, rewire edges only, groups per warp
for parts in 1 2 3; do for gpw in 4; do
  echo "parts=$parts gpw=$gpw: $(RTPP_TICK_PARTS=$parts RTPP_TICK_GPW=$gpw python tools/tick_bench.py 20 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["tick_window_ms_flushed"], d["tick_window_ms_warm"])')"
done; done
for gpw in 1 2 8; do
  echo "parts=2 gpw=$gpw: $(RTPP_TICK_PARTS=2 RTPP_TICK_GPW=$gpw python tools/tick_bench.py 20 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["tick_window_ms_flushed"], d["tick_window_ms_warm"])')"
done

python -m pytest tests/test_engine_gpu.py tests/test_am_engine_gpu.py tests/test_planners_gpu.py -x -q 2>&1 | tail -3
python - <<'PY'
import sys, os, json
sys.path.insert(0, '.'); sys.path.insert(0, 'tools')
import numpy as np, rtpp_b200, bench
with rtpp_b200.Ctx(0) as ctx:
    r = bench.measure_replanning(ctx)
    print(json.dumps(r))
PY

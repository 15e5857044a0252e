#!/bin/bash
# Runs on a 1-GPU box: the whole GPU test tier, smoke(), and both bench arms; logs and JSON lines into gpurun_out/.
set -u
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=10 > $O/final_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -4 $O/final_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/final_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/final_bench.json 2> $O/final_bench.err; echo "bench rc=$?"; head -c 700 $O/final_bench.json; echo
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/final_bench_reference.json 2> $O/final_bench_reference.err; echo "reference rc=$?"

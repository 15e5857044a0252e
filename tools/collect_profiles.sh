#!/bin/bash
# Runs on a B200 box (gpurun): the round's measured evidence into gpurun_out/ (copied to profiles/ afterwards).
#   bench line (N = 1), reference arm, ncu launch list of the bench without extras, ncu --set full of the tick kernel and of the
#   explicit-list kernel, nvidia-smi clocks during the bench.
set -u
O=gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $O/r02_clocks.csv &
SMI=$!
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02_bench.json 2> $O/r02_bench.err
echo "bench rc=$?"
kill $SMI
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err
echo "reference rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras > $O/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_bench_launches.csv python bench.py --steps 20 --warmup 5 --no-extras > $O/ncu_launches.log 2>&1
echo "launch list rc=$?"
timeout 300 python tools/tick_bench.py 3 > $O/plain2.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tick_check_kernel|edge_check_f32_kernel" -s 4 -c 3 -o $O/r02_k1w_full python tools/tick_bench.py 3 > $O/ncu_full.log 2>&1
echo "full rc=$?"
ls -la $O/r02_*

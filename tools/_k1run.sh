python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/s2_bench_v4.json 2> gpurun_out/s2_bench_v4.err; echo "bench exit $?"
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/s2_bench_v4_reference.json 2>/dev/null; echo "ref exit $?"
python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/plain4.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v5.csv python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/ncu5a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:edge_check_f32 -s 5 -c 2 -f -o gpurun_out/prof_k1_v5 python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/ncu5b.log 2>&1; echo done

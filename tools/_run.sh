S=$(date +%s)
python bench.py > gpurun_out/s2_bench_v6.json 2> gpurun_out/s2_bench_v6.err; echo "bench exit $? in $(( $(date +%s) - S )) s"

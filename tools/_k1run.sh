python -m pytest tests/test_edges_gpu.py -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 300 --warmup 10 --no-extras 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value']/1e9, d['roofline']['frac'], d['e2e']['value']/1e9, d['e2e']['matches_resident_run'])"

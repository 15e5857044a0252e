python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 300 --warmup 10 --no-extras 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value']/1e9, d['roofline']['frac'], d['e2e'])"

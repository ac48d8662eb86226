timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 120 python tools/prof_ops.py attn 2 2>&1 | grep flash
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'])"

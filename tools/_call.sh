# Scratch command list for one `gpurun -- 'bash tools/_call.sh'` round trip: GPU parity tests, then a short bench.
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['last_total_loss'])"

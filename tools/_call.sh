timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "presplit or split_bf16 or gemm_pair" 2>&1 | tail -8
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
for ws in 0 1; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --set gemm_wsplit=$ws 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('wsplit $ws', d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['last_total_loss'])"
done

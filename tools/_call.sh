timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for v in 0 1; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --set gemm_tiny_simt=$v 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tiny $v', d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['last_total_loss'])"
done

timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for cfg in "wgrad_overlap=0" "wgrad_units=1" "wgrad_units=2" "wgrad_units=4"; do
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --set $cfg 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$cfg', d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['last_total_loss'])"
done

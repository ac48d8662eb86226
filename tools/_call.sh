timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "layernorm" 2>&1 | tail -3
for bn in 128 256; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --set gemm_bn128_max=$bn 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bn128max $bn', d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['last_total_loss'])"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/t_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/t_ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/t_launches.csv 2>/dev/null | head -26

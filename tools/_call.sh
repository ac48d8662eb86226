timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r1_tests9.log; cat gpurun_out/r1_tests9.log
timeout 300 python tools/gpu_diag.py --cfg cfg2 --backend 0 --batch 8 --precision 2 --quiet 2>&1 | tail -5
timeout 300 python tools/gpu_diag.py --cfg cfg2 --backend 0 --batch 8 --precision 1 --quiet 2>&1 | tail -5
for p in 1 2; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision $p 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('precision $p', d['ms_per_step'], d['value'], d['host_submit_ms_per_step'], d['gpu_launches'])"; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 350 --csv --log-file gpurun_out/r1_launches_p2b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --precision 2 > /dev/null 2>&1

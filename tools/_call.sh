timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -x 2>&1 | tail -5
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r1_bench2.json 2> gpurun_out/r1_bench2.err; cat gpurun_out/r1_bench2.json; tail -5 gpurun_out/r1_bench2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 300 --csv --log-file gpurun_out/r1_launches_graph.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1_ncu_bench3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm2_k -s 2 -c 1 -o gpurun_out/r1_gemm2 python tools/prof_ops.py gemm 1 > gpurun_out/r1_ncu_gemm2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn2_ -s 3 -c 3 -o gpurun_out/r1_attn2 python tools/prof_ops.py attn 1 > gpurun_out/r1_ncu_attn2.log 2>&1
ls -la gpurun_out | tail -8

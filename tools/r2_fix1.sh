#!/bin/bash
TAG=${1:-r2_x1}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_benched_path.py -m gpu -q -k "whole_row or short or graph_replay" 2>&1 | tail -12 | tee $OUT/${TAG}_pytest.log
timeout 200 python tools/short_diag.py eng 2>&1 | tail -12 | tee $OUT/${TAG}_short_eng.log
for i in $(seq 16); do timeout 60 python tools/loader_first_use.py 2>&1 | tail -1; done | tee $OUT/${TAG}_first_use.log
timeout 600 python bench.py --steps 30 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 1200 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err

#!/bin/bash
# gpurun --timeout 600 -- 'bash tools/r2_ab5.sh TAG': pair-GEMM epilogue with coalesced global stores (gemm_direct_store)
TAG=${1:-r2_ab5}
OUT=gpurun_out; mkdir -p $OUT
SKF_OPTIONS=gemm_direct_store=1 timeout 400 python -m pytest tests/test_gpu_ops.py tests/test_gpu_benched_path.py -m gpu -q -x -k "gemm or benched or sublayer" 2>&1 | tail -4 | tee $OUT/${TAG}_pytest.log
for o in 0 1; do
  SKF_OPTIONS=gemm_direct_store=$o timeout 120 python tools/gemm_ws_time.py 2>&1 | sed "s/^/direct_store=$o /"
done | tee $OUT/${TAG}_gemm.log
for o in 0 1 0 1; do
  timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-loader --set gemm_direct_store=$o > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
  python -c "import json,sys; d=json.load(open('$OUT/${TAG}_bench.json')); print('direct_store=$o', round(d['ms_per_step'],4), round(d['value'],1), round(d['e2e']['value'],1), d['e2e']['last_total_loss'], round(d['roofline']['us_per_launch'],2))" || tail -5 $OUT/${TAG}_bench.err
done 2>&1 | tee $OUT/${TAG}_ab.log

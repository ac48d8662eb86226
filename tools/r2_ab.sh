#!/bin/bash
# gpurun --timeout 900 -- 'bash tools/r2_ab.sh TAG': programmatic dependent launch on / off (parity of the benchmarked path
# and the eager train tests with it on, then the bench line both ways), the fixed cost of a step (batch 128 vs 256), and
# the no-A-split timing build of the pair GEMM (upper bound of what activation images would buy).
TAG=${1:-r2_ab1}
OUT=gpurun_out; mkdir -p $OUT
SKF_OPTIONS=pdl=1 timeout 600 python -m pytest tests/test_gpu_benched_path.py tests/test_gpu_train.py tests/test_gpu_ops.py tests/test_gpu_decode.py -m gpu -q -x 2>&1 | tail -6 | tee $OUT/${TAG}_pytest_pdl.log
for p in 0 1 0 1; do
  timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-loader --set pdl=$p > $OUT/${TAG}_bench_pdl$p.json 2> $OUT/${TAG}_bench_pdl$p.err
  python -c "import json,sys; d=json.load(open('$OUT/${TAG}_bench_pdl$p.json')); print('pdl=$p', d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['last_total_loss'], d['roofline']['us_per_launch'])" || tail -5 $OUT/${TAG}_bench_pdl$p.err
done 2>&1 | tee $OUT/${TAG}_pdl.log
for b in 128 256; do
  timeout 200 python tools/time_cfg.py cfg2 $b
  SKF_OPTIONS=pdl=1 timeout 200 python tools/time_cfg.py cfg2 $b
done 2>&1 | tee $OUT/${TAG}_fixed_cost.log
timeout 200 python tools/prof_ops.py gemm 3 2>&1 | grep -v "^done" | tee $OUT/${TAG}_gemm_product.log
SKF_LIB=$PWD/sketchformer_b200/libskf_noasplit.so timeout 200 python tools/prof_ops.py gemm 3 2>&1 | tee $OUT/${TAG}_gemm_noasplit.log

#!/bin/bash
# gpurun --timeout 900 -- 'bash tools/r2_ab3.sh TAG': pair-GEMM tail split (parity, then the bench line with it off / on),
# LayerNorm backward with 4 rows per warp
TAG=${1:-r2_ab3}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_benched_path.py tests/test_gpu_train.py -m gpu -q -x -k "gemm or layernorm or benched or train" 2>&1 | tail -6 | tee $OUT/${TAG}_pytest.log
timeout 200 python tools/ln_ab.py 2>&1 | tee $OUT/${TAG}_ln_ab.log
for o in "gemm_tail_split=0" "gemm_tail_split=1" "ln_bwd_rows=4" "gemm_tail_split=0"; do
  timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-loader --set $o > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
  python -c "import json,sys; d=json.load(open('$OUT/${TAG}_bench.json')); print('$o', round(d['ms_per_step'],4), round(d['value'],1), round(d['e2e']['value'],1), d['e2e']['last_total_loss'], round(d['roofline']['us_per_launch'],2))" || tail -5 $OUT/${TAG}_bench.err
done 2>&1 | tee $OUT/${TAG}_ab.log
SKF_OPTIONS=gemm_tail_split=0 timeout 200 python tools/prof_ops.py gemm 3 2>&1 | grep -v "^done" | tee $OUT/${TAG}_gemm_nosplit.log
timeout 200 python tools/prof_ops.py gemm 3 2>&1 | grep -v "^done" | tee $OUT/${TAG}_gemm_tailsplit.log

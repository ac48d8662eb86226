#!/bin/bash
# gpurun --timeout 600 -- 'bash tools/r2_ab6.sh TAG': attention forward without the row-maximum pass (attn3_fastmax)
TAG=${1:-r2_ab6}
OUT=gpurun_out; mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_ops.py tests/test_gpu_benched_path.py tests/test_gpu_train.py -m gpu -q -x -k "attention or benched or train_step_matches" 2>&1 | tail -4 | tee $OUT/${TAG}_pytest.log
for o in 0 1; do
  SKF_OPTIONS=attn3_fastmax=$o timeout 120 python tools/attn_time.py 7 2>&1 | sed "s/^/fastmax=$o /"
done | tee $OUT/${TAG}_attn.log
for o in 0 1 0 1; do
  timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-loader --set attn3_fastmax=$o > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
  python -c "import json,sys; d=json.load(open('$OUT/${TAG}_bench.json')); print('fastmax=$o', round(d['ms_per_step'],4), round(d['value'],1), round(d['e2e']['value'],1), d['e2e']['last_total_loss'])" || tail -5 $OUT/${TAG}_bench.err
done 2>&1 | tee $OUT/${TAG}_ab.log

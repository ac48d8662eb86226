#!/bin/bash
# gpurun --timeout 600 -- 'bash tools/r2_race.sh TAG [tc]': compute-sanitizer racecheck on the SIMT / cluster kernels' op
# tests (tc: on a few tcgen05 GEMM / attention cases instead)
TAG=${1:-r2_race}
OUT=gpurun_out; mkdir -p $OUT
SEL="pool_fwd_bwd or cross_attention_on_the_rank2 or layernorm_fwd_bwd or colsum or kv_expand or token_and_class_ce or embed"
[ "$2" = "tc" ] && SEL="(pair_presplit and 1600) or (attention_tcgen05_whole_row_with_images and 2-8-200-200-32-1) or (gemm_tcgen05_3xtf32 and 333) or (attention_tcgen05 and 3-8-200-200-32-1) or decode"
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 10 \
  python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "$SEL" > $OUT/${TAG}_racecheck.log 2>&1
echo "exit $?" >> $OUT/${TAG}_racecheck.log
grep -E "passed|failed|RACECHECK SUMMARY|exit|Potential|hazard" $OUT/${TAG}_racecheck.log | sort | uniq -c | head -20

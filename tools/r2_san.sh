#!/bin/bash
# gpurun --timeout 900 -- 'bash tools/r2_san.sh TAG': compute-sanitizer over the GPU tests (memcheck on the train / decode /
# loader / op suites at their small sizes, synccheck on a few tcgen05 ops)
TAG=${1:-r2_san}
OUT=gpurun_out; mkdir -p $OUT
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_train.py tests/test_gpu_decode.py tests/test_loader.py tests/test_evaluation.py tests/test_gpu_reference_golden.py -m gpu -q -x \
  -k "not cfg5 and not cfg2-b6" > $OUT/${TAG}_memcheck_train.log 2>&1
echo "exit $?" >> $OUT/${TAG}_memcheck_train.log
tail -6 $OUT/${TAG}_memcheck_train.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "not 25600 and not 25472 and not 40000 and not 19100 and not many_items" > $OUT/${TAG}_memcheck_ops.log 2>&1
echo "exit $?" >> $OUT/${TAG}_memcheck_ops.log
tail -6 $OUT/${TAG}_memcheck_ops.log
timeout 300 compute-sanitizer --tool synccheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "(pair_presplit and 1600) or (attention_tcgen05_whole_row_with_images and 2-8-200-200-32-1) or (tail_split and 19200 and fwd)" > $OUT/${TAG}_synccheck.log 2>&1
echo "exit $?" >> $OUT/${TAG}_synccheck.log
tail -6 $OUT/${TAG}_synccheck.log

#!/bin/bash
# gpurun --timeout 600 -- 'bash tools/r2_t7.sh TAG': decode tests (incl. the uncached growing-mask mode)
TAG=${1:-r2_t7}
OUT=gpurun_out; mkdir -p $OUT
timeout 500 python -m pytest tests/test_gpu_decode.py -m gpu -q -x 2>&1 | tail -8 | tee $OUT/${TAG}_pytest.log

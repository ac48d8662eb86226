#!/bin/bash
# gpurun --timeout 600 -- 'bash tools/r2_ab4.sh TAG': what bounds the pair GEMM — timing builds with parts switched off
# (results wrong, timing valid), and the new sublayer entry-point tests
TAG=${1:-r2_ab4}
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "sublayer" 2>&1 | tail -5 | tee $OUT/${TAG}_pytest.log
for v in "" _noasplit _noaload _noepi _noepi_noaload _st3 _st2; do
  SKF_LIB=$PWD/sketchformer_b200/libskf$v.so timeout 120 python tools/gemm_ws_time.py 2>&1 | tail -6
done | tee $OUT/${TAG}_gemm_variants.log

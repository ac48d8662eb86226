#!/bin/bash
# gpurun -- 'bash tools/r2_attn.sh TAG': op-level parity of the whole-row attention kernels + timings + phase traces
TAG=${1:-r2_a1}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "whole_row" 2>&1 | tail -12 | tee $OUT/${TAG}_pytest.log
timeout 300 python tools/attn_time.py 6 7 2>&1 | tee $OUT/${TAG}_attn_time.log
for w in fwd dq dkv; do timeout 120 python tools/attn_trace.py 200 1 $w > $OUT/${TAG}_trace_$w.log 2>&1; done

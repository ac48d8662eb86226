#!/bin/bash
# gpurun -- 'bash tools/r2_probe.sh': tcgen05 layout / timing probe, then the GPU test-suite (no -x: all failures listed)
OUT=gpurun_out
mkdir -p $OUT
for v in 1 2 4; do timeout 60 tools/probe/mma_probe $v 2>&1 | tee -a $OUT/r2_mma_probe.log; done
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee $OUT/r2_s2_pytest.log

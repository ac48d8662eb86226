#!/bin/bash
TAG=${1:-r2_t8}
OUT=gpurun_out; mkdir -p $OUT
timeout 800 python -m pytest tests/test_gpu_benched_path.py -m gpu -q -x 2>&1 | tail -12 | tee $OUT/${TAG}_pytest.log

#!/bin/bash
TAG=${1:-r2_ab9}
OUT=gpurun_out; mkdir -p $OUT
for v in "" _noconv; do
  SKF_LIB=$PWD/sketchformer_b200/libskf$v.so timeout 120 python tools/attn_time.py 6 7 2>&1 | grep fwd | sed "s/^/libskf$v /"
done | tee $OUT/${TAG}_attn_noconv.log

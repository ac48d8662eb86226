#!/bin/bash
# gpurun -- 'bash tools/r2_ncu_attn.sh TAG REGEX': ncu --set full of the attention kernels matching REGEX
TAG=${1:-r2_a2}
RE=${2:-attn3_}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$RE -s 6 -c 2 -o $OUT/${TAG}_attn3 -f \
  python tools/attn_time.py 6 > $OUT/${TAG}_ncu.log 2>&1
tail -3 $OUT/${TAG}_ncu.log
ls -la $OUT/${TAG}_attn3.ncu-rep

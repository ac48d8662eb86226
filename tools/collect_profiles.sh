#!/bin/bash
# Run on the GPU box (gpurun -- 'bash tools/collect_profiles.sh TAG'): bench line, per-launch list of one train
# step, and `ncu --set full` captures of the dominant kernels; everything lands in gpurun_out/ and the text
# summaries are copied into profiles/ by hand afterwards.
TAG=${1:-r1_final}
OUT=gpurun_out
mkdir -p $OUT
python bench.py --steps 50 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 400 $OUT/${TAG}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
python tools/launch_summary.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_launches.summary.txt
head -30 $OUT/${TAG}_launches.summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm3_k -s 2 -c 1 -o $OUT/${TAG}_gemm3 -f \
  python tools/prof_ops.py gemm 2 > $OUT/${TAG}_ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn2_ -s 6 -c 3 -o $OUT/${TAG}_attn2 -f \
  python tools/prof_ops.py attn 2 > $OUT/${TAG}_ncu_attn.log 2>&1
python tools/ncu_summary.py $OUT/${TAG}_gemm3.ncu-rep > $OUT/${TAG}_ncu_full_gemm3_pair_ffn1.txt
python tools/ncu_summary.py $OUT/${TAG}_attn2.ncu-rep > $OUT/${TAG}_ncu_full_attn2.txt
ls -la $OUT | tail -12

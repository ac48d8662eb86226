#!/bin/bash
# gpurun --timeout 1200 -- 'bash tools/r2_measure.sh TAG': GPU parity suite, smoke, bench line, per-launch list of one train
# step (device time + DRAM bytes of every kernel), and `ncu --set full` captures of the attention and pair-GEMM kernels.
TAG=${1:-r2_m1}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee $OUT/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.log
timeout 600 python bench.py --steps 30 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 600 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv \
  --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
python tools/launch_summary.py $OUT/${TAG}_launches.csv 40 > $OUT/${TAG}_launches.summary.txt
python tools/hbm_summary.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_hbm.summary.txt
head -45 $OUT/${TAG}_launches.summary.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn3_ -s 6 -c 3 -o $OUT/${TAG}_attn3 -f \
  python tools/attn_time.py 7 > $OUT/${TAG}_ncu_attn.log 2>&1
tail -4 $OUT/${TAG}_ncu_attn.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm3_k -s 2 -c 1 -o $OUT/${TAG}_gemm3 -f \
  python tools/prof_ops.py gemm 2 > $OUT/${TAG}_ncu_gemm.log 2>&1
python tools/ncu_summary.py $OUT/${TAG}_attn3.ncu-rep > $OUT/${TAG}_ncu_full_attn3.txt
python tools/ncu_summary.py $OUT/${TAG}_gemm3.ncu-rep > $OUT/${TAG}_ncu_full_gemm3_pair_ffn1.txt
timeout 120 python tools/attn_time.py 2 7 2>&1 | tee $OUT/${TAG}_attn_time.log
ls -la $OUT | tail -14

#!/bin/bash
# One gpurun round trip (round 2): GPU parity tests, smoke, bench line, and a per-launch list of one train step with
# device time and DRAM bytes of EVERY kernel (ncu, two cheap metric passes).  Everything lands in gpurun_out/.
#   gpurun --timeout 900 -- 'bash tools/r2_call.sh TAG [tests|notests]'
TAG=${1:-r2_s1}
OUT=gpurun_out
mkdir -p $OUT
if [ "${2:-tests}" = "tests" ]; then
  timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee $OUT/${TAG}_pytest.log
  timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.log
fi
timeout 600 python bench.py --steps 30 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 600 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv \
  --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
python tools/launch_summary.py $OUT/${TAG}_launches.csv 40 > $OUT/${TAG}_launches.summary.txt
python tools/hbm_summary.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_hbm.summary.txt
head -45 $OUT/${TAG}_launches.summary.txt

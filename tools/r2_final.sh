#!/bin/bash
# gpurun --timeout 1500 -- 'bash tools/r2_final.sh TAG': the whole round-end evidence set in one call —
# GPU parity suite, smoke, bench line (cfg2), launch list of one train step with DRAM bytes per kernel,
# `ncu --set full` of the attention forward / backward kernels and the pair GEMM, bench lines of the other rows
# of SURVEY.md 8(d) (cfg4, fwd, decode, cfg5).
TAG=${1:-r2_f1}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee $OUT/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.log
timeout 600 python bench.py --steps 50 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 600 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv \
  --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-loader > $OUT/${TAG}_ncu_bench.log 2>&1
python tools/launch_summary.py $OUT/${TAG}_launches.csv 40 > $OUT/${TAG}_launches.summary.txt
python tools/hbm_summary.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_hbm.summary.txt
head -24 $OUT/${TAG}_launches.summary.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn3_fwd -s 3 -c 1 -o $OUT/${TAG}_attn3_fwd -f \
  python tools/attn_time.py 7 > $OUT/${TAG}_ncu_attn_fwd.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn3_bwd -s 6 -c 2 -o $OUT/${TAG}_attn3_bwd -f \
  python tools/attn_time.py 7 > $OUT/${TAG}_ncu_attn_bwd.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm3_k -s 2 -c 1 -o $OUT/${TAG}_gemm3 -f \
  python tools/prof_ops.py gemm 2 > $OUT/${TAG}_ncu_gemm.log 2>&1
{ python tools/ncu_summary.py $OUT/${TAG}_attn3_fwd.ncu-rep; python tools/ncu_summary.py $OUT/${TAG}_attn3_bwd.ncu-rep; } > $OUT/${TAG}_ncu_full_attn3.txt
python tools/ncu_summary.py $OUT/${TAG}_gemm3.ncu-rep > $OUT/${TAG}_ncu_full_gemm3_pair_ffn1.txt
timeout 120 python tools/attn_time.py 2 7 2>&1 | tee $OUT/${TAG}_attn_time.log
for cfg in cfg4 fwd decode cfg5; do
  timeout 600 python bench.py --config $cfg --steps 20 --warmup 3 > $OUT/${TAG}_bench_$cfg.json 2> $OUT/${TAG}_bench_$cfg.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_$cfg.json").read().strip().splitlines()[-1])
    print("$cfg", round(d["value"], 1), "sk/s", round(d["ms_per_step"], 3), "ms e2e", round(d["e2e"]["value"], 1), "cpu", d.get("cpu_baseline", {}).get("value"), "tok/s", d.get("tokens_per_s"), "roof", round(d["roofline"]["frac"], 4))
except Exception as e:
    print("$cfg FAILED", e); print(open("$OUT/${TAG}_bench_$cfg.err").read()[-1500:])
PY
done 2>&1 | tee $OUT/${TAG}_cfgs.log
ls -la $OUT | tail -30

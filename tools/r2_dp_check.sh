#!/bin/bash
# gpurun --gpus N -- 'bash tools/r2_dp_check.sh TAG N': the driver's own multi-GPU command (default flags) and the strong-scaling line
TAG=${1:-r2_dpc}
N=${2:-2}
OUT=gpurun_out; mkdir -p $OUT
for mode in ${3:-weak strong}; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 20 --warmup 5 --scaling $mode --no-loader > $OUT/${TAG}_${N}_$mode.json 2> $OUT/${TAG}_${N}_$mode.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_${N}_$mode.json").read().strip().splitlines()[-1])
    print("$mode N=$N", round(d["value"], 1), "sk/s", round(d["ms_per_step"], 3), "ms e2e", round(d["e2e"]["value"], 1), "dp_err", d.get("dp_grad_rel_err"), d["run"]["dp_exchange"][:50])
except Exception as e:
    print("$mode FAILED", e); print(open("$OUT/${TAG}_${N}_$mode.err").read()[-2000:])
PY
done 2>&1 | tee $OUT/${TAG}_${N}.log

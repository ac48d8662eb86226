#!/bin/bash
# gpurun --gpus 8 -- 'bash tools/r2_dp8.sh TAG': 8-rank weak scaling with the exchange INSIDE the step (4 buckets, graph-captured)
TAG=${1:-r2_dp8b}
OUT=gpurun_out; mkdir -p $OUT
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 \
  bench.py --gpus 8 --steps 20 --warmup 3 --dp-buckets 4 --no-loader --no-cpu-baseline > $OUT/${TAG}_weak_b4.json 2> $OUT/${TAG}_weak_b4.err
echo "rc $?"
python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_weak_b4.json").read().strip().splitlines()[-1])
    print("weak in-step N=8", round(d["value"], 1), "sk/s", round(d["ms_per_step"], 3), "ms e2e", round(d["e2e"]["value"], 1), "dp_err", d.get("dp_grad_rel_err"), d["run"]["dp_exchange"][:50])
except Exception as e:
    print("FAILED", e); print(open("$OUT/${TAG}_weak_b4.err").read()[-1500:])
PY

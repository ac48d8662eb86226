#!/bin/bash
# gpurun --gpus N -- 'bash tools/r2_dp.sh TAG N': data-parallel bench lines (weak, bucketed in-graph exchange vs one torch
# all_reduce after the step; strong scaling), each with the dp gradient check
TAG=${1:-r2_dp}
N=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
run() {  # name, extra args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 30 --warmup 5 $2 > $OUT/${TAG}_$1.json 2> $OUT/${TAG}_$1.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_$1.json").read().strip().splitlines()[-1])
    print("$1", "N=$N", round(d["value"], 1), "sk/s", round(d["ms_per_step"], 3), "ms", "e2e", round(d["e2e"]["value"], 1), "dp_err", d.get("dp_grad_rel_err"), d["run"]["dp_exchange"][:40])
except Exception as e:
    print("$1 FAILED", e); print(open("$OUT/${TAG}_$1.err").read()[-2500:])
PY
}
run weak_b4 "--dp-buckets 4"
#run weak_b1 "--dp-buckets 1"
run weak_torch "--dp-buckets 0"
run strong_b4 "--scaling strong --dp-buckets 4"
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_single.json 2> $OUT/${TAG}_single.err
python -c "
import json; d=json.loads(open('$OUT/${TAG}_single.json').read().strip().splitlines()[-1]); print('single', round(d['value'],1), round(d['ms_per_step'],3))"

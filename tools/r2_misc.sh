#!/bin/bash
TAG=${1:-r2_m1}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_loader.py tests/test_gpu_decode.py -m gpu -q 2>&1 | tail -8 | tee $OUT/${TAG}_pytest.log
for cfg in cfg2 decode; do
  timeout 600 python bench.py --config $cfg --steps 20 --warmup 3 > $OUT/${TAG}_bench_$cfg.json 2> $OUT/${TAG}_bench_$cfg.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_$cfg.json").read().strip().splitlines()[-1])
    print("$cfg", round(d["value"], 1), "sk/s", round(d["ms_per_step"], 3), "ms e2e", round(d["e2e"]["value"], 1), "tok/s", d.get("tokens_per_s"), "launches", d["gpu_launches"])
    print("   loader:", d.get("e2e_loader"))
except Exception as e:
    print("$cfg FAILED", e); print(open("$OUT/${TAG}_bench_$cfg.err").read()[-2500:])
PY
done

#!/bin/bash
# gpurun -- 'bash tools/r2_cfgs.sh TAG': GPU tests, then one bench line per measurement row of SURVEY.md 8(d)
TAG=${1:-r2_s3}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee $OUT/${TAG}_pytest.log
for cfg in cfg2 cfg4 fwd decode cfg5; do
  timeout 900 python bench.py --config $cfg --steps 20 --warmup 3 > $OUT/${TAG}_bench_$cfg.json 2> $OUT/${TAG}_bench_$cfg.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_$cfg.json").read().strip().splitlines()[-1])
    print("$cfg", round(d["value"], 1), "sk/s", round(d["ms_per_step"], 3), "ms e2e", round(d["e2e"]["value"], 1), "cpu", d.get("cpu_baseline", {}).get("value"), "tok/s", d.get("tokens_per_s"), "roof", round(d["roofline"]["frac"], 4))
except Exception as e:
    print("$cfg FAILED", e); print(open("$OUT/${TAG}_bench_$cfg.err").read()[-1500:])
PY
done

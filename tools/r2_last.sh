#!/bin/bash
TAG=${1:-r2_last}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee $OUT/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee $OUT/${TAG}_smoke.log
timeout 600 python bench.py --steps 50 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python -c "import json; d=json.load(open('$OUT/${TAG}_bench.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'])"

#!/bin/bash
# gpurun -- 'bash tools/r2_full.sh TAG': whole GPU suite (no -x), the loader test repeated, smoke, default bench line
TAG=${1:-r2_f1}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee $OUT/${TAG}_pytest.log
for i in 1 2 3 4 5; do timeout 120 python -m pytest tests/test_loader.py -m gpu -q 2>&1 | tail -4; done | tee $OUT/${TAG}_loader_rep.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.log
timeout 600 python bench.py --steps 30 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 900 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err

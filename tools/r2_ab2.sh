#!/bin/bash
# gpurun --timeout 600 -- 'bash tools/r2_ab2.sh TAG': bench line per programmatic-dependent-launch class mask
# (1 SIMT kernels, 2 tcgen05 GEMMs, 4 tcgen05 attention, 8 late trigger)
TAG=${1:-r2_ab2}
OUT=gpurun_out; mkdir -p $OUT
for p in 0 1 6 14 15 7 9 0; do
  timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-loader --set pdl=$p > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
  python -c "import json,sys; d=json.load(open('$OUT/${TAG}_bench.json')); print('pdl=$p', round(d['ms_per_step'],4), round(d['value'],1), round(d['e2e']['value'],1), d['e2e']['last_total_loss'], round(d['roofline']['us_per_launch'],2))" || tail -5 $OUT/${TAG}_bench.err
done 2>&1 | tee $OUT/${TAG}_pdl.log

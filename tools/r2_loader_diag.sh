#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python tools/loader_diag.py 2>&1 | tail -60 | tee $OUT/loader_diag.log
timeout 240 compute-sanitizer --tool memcheck --print-limit 5 python tools/loader_diag.py quick 2>&1 | tail -40 | tee $OUT/loader_diag_memcheck.log
timeout 240 compute-sanitizer --tool racecheck --print-limit 5 python tools/loader_diag.py quick 2>&1 | tail -40 | tee $OUT/loader_diag_racecheck.log

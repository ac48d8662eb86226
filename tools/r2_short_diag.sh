#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python tools/short_diag.py op 2>&1 | tail -70 | tee $OUT/short_diag_op.log
timeout 200 python tools/short_diag.py eng 2>&1 | tail -15 | tee $OUT/short_diag_eng.log
timeout 200 python tools/short_diag.py eng attn3=0 2>&1 | tail -15 | tee $OUT/short_diag_eng_noattn3.log

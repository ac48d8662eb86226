timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r1_tests5.log; cat gpurun_out/r1_tests5.log
for cfg in "0 1" "1 1" "2 1" "1 0" "2 0"; do set -- $cfg; echo "== precision $1 attn_backend $2"; timeout 300 python tools/gpu_diag.py --cfg cfg2 --backend 0 --batch 8 --precision $1 --attn-backend $2 --quiet 2>&1 | tail -5; done > gpurun_out/r1_diag_matrix.log 2>&1
cat gpurun_out/r1_diag_matrix.log

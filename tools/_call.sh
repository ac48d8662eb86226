timeout 200 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "embed or layernorm" 2>&1 | tail -2
timeout 200 python tools/ln_ab.py 2>&1 | tail -8

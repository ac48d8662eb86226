"""Times the pair GEMM as the plan's forward / dgrad products run it (bf16x3, weights pre-split into bf16 images) at the
cfg2 layer shapes, M = 25600: CUDA events over 20 launches, outputs cycling over more than L2.
    [SKF_LIB=...variant.so] python tools/gemm_ws_time.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sketchformer_b200.ops import gemm, split_bf16  # noqa: E402

M = 25600
torch.manual_seed(0)
tag = os.path.basename(os.environ.get("SKF_LIB", "libskf.so"))
for name, K, N, fwd in (("FFN-1 fwd", 256, 1024, True), ("FFN-2 fwd", 1024, 256, True), ("QKV fwd", 256, 768, True),
                        ("WO fwd", 256, 256, True), ("FFN-1 dgrad", 1024, 256, False), ("FFN-2 dgrad", 256, 1024, False)):
    X = torch.randn(M, K, device="cuda")
    W = torch.randn(K, N, device="cuda") if fwd else torch.randn(N, K, device="cuda")     # dgrad: W stored [out=N.., in]
    bias = torch.zeros(N, device="cuda")
    nbuf = max(2, min(4, int(1e9 // (M * N * 4))))
    Ys = [torch.empty(M, N, device="cuda") for _ in range(nbuf)]
    hi, lo = split_bf16(W)
    sb = (N, 1) if fwd else (1, K)

    def run(i):
        gemm(X, K, 1, W, sb[0], sb[1], Ys[i % nbuf], N, M, N, K, 5, bias=bias, act=1 if fwd else 0, b_hi16=hi, b_lo16=lo)
    for i in range(3):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    print(f"{tag:28s} {name:12s} {M}x{K}x{N}: {us:6.1f} us  {2.0 * M * N * K / us / 1e6:6.1f} TFLOP/s", flush=True)

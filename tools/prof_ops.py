"""Runs the dominant kernels of the cfg2 step in isolation (for `ncu --set full -k regex:...`):
FFN-1 GEMM (25600x256 . 256x1024, bias+ReLU), encoder self-attention fwd/bwd (B128 H8 T200 dh32).
    python tools/prof_ops.py [gemm|attn|all] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sketchformer_b200 import _lib                              # noqa: E402
from sketchformer_b200.ops import gemm, ptr, split_bf16, stream  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
lib = _lib.load()
torch.manual_seed(0)
if what in ("gemm", "all"):
    M, K, N = 25600, 256, 1024
    X = torch.randn(M, K, device="cuda")
    W = torch.randn(K, N, device="cuda")
    b = torch.zeros(N, device="cuda")
    Y = torch.empty(M, N, device="cuda")
    import time
    # the plan's forward products read the weight's pre-split bf16 images (gemm3_k<..., WS = true>)
    img = dict(zip(("b_hi16", "b_lo16"), split_bf16(W)))
    for _ in range(reps):
        gemm(X, K, 1, W, N, 1, Y, N, M, N, K, 5, bias=b, act=1, **img)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        gemm(X, K, 1, W, N, 1, Y, N, M, N, K, 5, bias=b, act=1, **img)
    e1.record()
    torch.cuda.synchronize()
    print(f"FFN-1 gemm backend 5 + weight images: {e0.elapsed_time(e1) * 1e3 / 20:.1f} us")
    for backend in (2, 5, 0):
        for _ in range(reps):
            gemm(X, K, 1, W, N, 1, Y, N, M, N, K, backend, bias=b, act=1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            gemm(X, K, 1, W, N, 1, Y, N, M, N, K, backend, bias=b, act=1)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 20
        print(f"FFN-1 gemm backend {backend}: {us:.1f} us  {2.0 * M * N * K / us / 1e6:.1f} TFLOP/s (algorithmic)")
    for (m, n, k, nm) in ((M, K, N, "FFN-2 fwd"), (M, 3 * K, K, "QKV fwd"), (M, K, K, "WO fwd")):
        Xa = torch.randn(m, k, device="cuda"); Wb = torch.randn(k, n, device="cuda"); Yc = torch.empty(m, n, device="cuda")
        for backend in (2, 5):
            for _ in range(reps):
                gemm(Xa, k, 1, Wb, n, 1, Yc, n, m, n, k, backend)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                gemm(Xa, k, 1, Wb, n, 1, Yc, n, m, n, k, backend)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / 20
            print(f"{nm} backend {backend}: {us:.1f} us  {2.0 * m * n * k / us / 1e6:.1f} TFLOP/s (algorithmic)")
        # dgrad (B K-major) and wgrad (both MN-major, split-K) of the same layer
        dX = torch.empty(m, k, device="cuda")
        dW = torch.zeros(k, n, device="cuda")
        for backend in (2, 5):
            for what2 in ("dgrad", "wgrad"):
                def run():
                    if what2 == "dgrad":
                        gemm(Yc, n, 1, Wb, 1, n, dX, k, m, k, n, backend)
                    else:
                        gemm(Xa, 1, k, Yc, n, 1, dW, n, k, n, m, backend, accumulate=1, split_k=max(1, min(m // 256, (74 if backend == 5 else 148) // max(1, ((k + 255) // 256 if backend == 5 else (k + 127) // 128) * ((n + 255) // 256)))))
                for _ in range(reps):
                    run()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(20):
                    run()
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) * 1e3 / 20
                print(f"{nm}->{what2} backend {backend}: {us:.1f} us  {2.0 * m * n * k / us / 1e6:.1f} TFLOP/s")
    # FFN-2 shape with residual
    W2 = torch.randn(N, K, device="cuda")
    Y2 = torch.empty(M, K, device="cuda")
    for _ in range(reps):
        gemm(Y, N, 1, W2, K, 1, Y2, K, M, K, N, 2, bias=b[:K].contiguous(), resid=X)
if what in ("attn", "all"):
    B, H, T, dh = 128, 8, 200, 32
    d = H * dh
    qkv = torch.randn(B * T, 3 * d, device="cuda")
    O = torch.empty(B * T, d, device="cuda")
    dO = torch.randn(B * T, d, device="cuda")
    dqkv = torch.empty_like(qkv)
    lse = torch.empty(B * H * T, device="cuda")
    delta = torch.empty_like(lse)
    pad = torch.zeros(B * T, dtype=torch.uint8, device="cuda")
    def run(fb, bb, mode=1):
        _lib.check(lib.skf_attn_fwd(ptr(qkv), 3 * d, ptr(qkv[:, d:]), 3 * d, ptr(qkv[:, 2 * d:]), 3 * d, ptr(O), d,
                                    ptr(lse), ptr(pad), B, H, T, T, dh, mode, 0, fb, stream()))
        if bb >= 0:
            _lib.check(lib.skf_attn_bwd(ptr(qkv), 3 * d, ptr(qkv[:, d:]), 3 * d, ptr(qkv[:, 2 * d:]), 3 * d, ptr(O), d,
                                        ptr(dO), d, ptr(lse), ptr(delta), ptr(dqkv), 3 * d, ptr(dqkv[:, d:]), 3 * d,
                                        ptr(dqkv[:, 2 * d:]), 3 * d, ptr(pad), B, H, T, T, dh, mode, 0, bb, stream()))

    for name, fb, bb in (("flash x3 fwd", 2, -1), ("flash x3 fwd+bwd", 2, 2), ("flash tf32 fwd", 0, -1),
                         ("flash tf32 fwd+bwd", 0, 0)):
        for mode in (1, 2):
            for _ in range(reps):
                run(fb, bb, mode)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                run(fb, bb, mode)
            e1.record()
            torch.cuda.synchronize()
            print(f"attention {name:22s} mode {mode}: {e0.elapsed_time(e1) * 100:.1f} us")
torch.cuda.synchronize()
print("done")

"""Times the pair GEMM on the cfg2 layer shapes as the plan runs them: forward and dgrad with the weight's pre-split
bf16 images (WS variant), wgrad with both operands split in the kernel.  CUDA events, 20 launches after 5 warm-ups.
    python tools/ws_ab.py [tag]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sketchformer_b200 import _lib                              # noqa: E402
from sketchformer_b200.ops import gemm, split_bf16              # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else ""
lib = _lib.load()
torch.manual_seed(0)
M = 25600


def timed(fn):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / 20


def rel(y, want):
    return float((y.double() - want).norm() / want.norm())


tot = 0.0
for (n, k, nm) in ((1024, 256, "FFN-1"), (256, 1024, "FFN-2"), (768, 256, "QKV"), (256, 256, "WO")):
    X = torch.randn(M, k, device="cuda")
    W = torch.randn(k, n, device="cuda") / k ** 0.5
    dY = torch.randn(M, n, device="cuda")
    Y = torch.empty(M, n, device="cuda")
    dX = torch.empty(M, k, device="cuda")
    dW = torch.zeros(k, n, device="cuda")
    hi, lo = split_bf16(W)
    sk = max(1, min(M // 256, 74 // max(1, ((k + 255) // 256) * ((n + 255) // 256))))
    t_f = timed(lambda: gemm(X, k, 1, W, n, 1, Y, n, M, n, k, 5, b_hi16=hi, b_lo16=lo))
    e_f = rel(Y, X.double() @ W.double())
    t_d = timed(lambda: gemm(dY, n, 1, W, 1, n, dX, k, M, k, n, 5, b_hi16=hi, b_lo16=lo))
    e_d = rel(dX, dY.double() @ W.double().T)
    t_w = timed(lambda: gemm(X, 1, k, dY, n, 1, dW, n, k, n, M, 5, accumulate=1, split_k=sk))
    tot += t_f + t_d + t_w
    print(f"{tag} {nm:6s} fwd {t_f:6.1f} us ({e_f:.1e}) | dgrad {t_d:6.1f} us ({e_d:.1e}) | wgrad {t_w:6.1f} us", flush=True)
print(f"{tag} sum {tot:.1f} us")

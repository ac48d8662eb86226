"""A/B of the pair GEMM's product schemes (skf_set_option gemm_mix = 0 / 1 / 2) on the cfg2 layer shapes:
time per launch (CUDA events, 20 launches after 5 warm-ups) and relative L2 error against the fp64 product.
    python tools/mix_ab.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sketchformer_b200 import _lib                              # noqa: E402
from sketchformer_b200.ops import gemm                          # noqa: E402

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


for (n, k, nm) in ((1024, 256, "FFN-1"), (256, 1024, "FFN-2"), (768, 256, "QKV"), (256, 256, "WO")):
    X = torch.randn(M, k, device="cuda")
    W = torch.randn(k, n, device="cuda") / k ** 0.5
    dY = torch.randn(M, n, device="cuda")
    Y = torch.empty(M, n, device="cuda")
    dX = torch.empty(M, k, device="cuda")
    dW = torch.zeros(k, n, device="cuda")
    sk = max(1, min(M // 256, 74 // max(1, ((k + 255) // 256) * ((n + 255) // 256))))
    Xd, Wd, dYd = X.double(), W.double(), dY.double()
    want = {"fwd": Xd @ Wd, "dgrad": dYd @ Wd.T, "wgrad": Xd.T @ dYd}
    for mix in (0, 1, 2):
        _lib.check(lib.skf_set_option(b"gemm_mix", mix))
        out = []
        t = timed(lambda: gemm(X, k, 1, W, n, 1, Y, n, M, n, k, 5))
        out.append(f"fwd {t:6.1f} us err {rel(Y, want['fwd']):.1e}")
        t = timed(lambda: gemm(dY, n, 1, W, 1, n, dX, k, M, k, n, 5))
        out.append(f"dgrad {t:6.1f} us err {rel(dX, want['dgrad']):.1e}")
        t = timed(lambda: gemm(X, 1, k, dY, n, 1, dW, n, k, n, M, 5, accumulate=1, split_k=sk))
        dW.zero_()
        gemm(X, 1, k, dY, n, 1, dW, n, k, n, M, 5, accumulate=1, split_k=sk)
        out.append(f"wgrad {t:6.1f} us err {rel(dW, want['wgrad']):.1e}")
        print(f"{nm:6s} [{M}x{k}x{n}] mix={mix}: " + " | ".join(out), flush=True)
_lib.check(lib.skf_set_option(b"gemm_mix", 1))

"""Times every distinct GEMM of the cfg2 train step (forward, dgrad, wgrad) through skf_gemm and prints
algorithmic TFLOP/s and the tensor-pipe time bound.  python tools/gemm_shapes.py [backend=2]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sketchformer_b200.ops import gemm  # noqa: E402

backend = int(sys.argv[1]) if len(sys.argv) > 1 else 2
Me, Md, d, f, V, u = 25600, 25472, 256, 1024, 1004, 256
passes = 3 if backend == 2 else 1
PEAK = 842.0e12   # TF32 burst (bf16 measured / 2)


def bench(tag, M, N, K, a_mn, b_mn, count, **kw):
    A = torch.randn((K, M) if a_mn else (M, K), device="cuda")
    B = torch.randn((K, N) if b_mn else (N, K), device="cuda")
    Y = torch.zeros(M, N, device="cuda")
    sa = (1, M) if a_mn else (K, 1)
    sb = (N, 1) if b_mn else (1, K)
    extra = {}
    if kw.get("split"):
        tiles = ((M + 127) // 128) * ((N + 255) // 256 if N > 128 else 1)
        extra = dict(accumulate=1, split_k=max(1, min(148 // tiles, K // 256)))
    for _ in range(3):
        gemm(A, sa[0], sa[1], B, sb[0], sb[1], Y, N, M, N, K, backend, **extra)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 10
    for _ in range(n):
        gemm(A, sa[0], sa[1], B, sb[0], sb[1], Y, N, M, N, K, backend, **extra)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    fl = 2.0 * M * N * K
    print(f"{tag:28s} M={M:6d} N={N:5d} K={K:6d} x{count:3d}  {us:7.1f} us  {fl / us / 1e6:6.1f} TF/s alg  "
          f"mma-bound {fl * passes / PEAK * 1e6:6.1f} us  total {us * count / 1e3:6.2f} ms")
    return us * count


tot = 0.0
# forward (A K-major activations, B = W[K,N] MN-major)
tot += bench("fwd qkv (enc+dec self)", Me, 3 * d, d, False, True, 8)
tot += bench("fwd wo / q2 (d->d)", Me, d, d, False, True, 16)
tot += bench("fwd kv2 (d->2d)", Me, 2 * d, d, False, True, 4)
tot += bench("fwd ffn1", Me, f, d, False, True, 8)
tot += bench("fwd ffn2", Me, d, f, False, True, 8)
tot += bench("fwd logits", Md, V, d, False, True, 1)
tot += bench("fwd pool", Me, u, d, False, True, 1)
# dgrad (A = dY K-major, B = W^T: K-major)
tot += bench("dgrad qkv", Me, d, 3 * d, False, False, 8)
tot += bench("dgrad d->d", Me, d, d, False, False, 17)
tot += bench("dgrad kv2", Me, d, 2 * d, False, False, 4)
tot += bench("dgrad ffn1", Me, d, f, False, False, 8)
tot += bench("dgrad ffn2", Me, f, d, False, False, 8)
tot += bench("dgrad logits", Md, d, V, False, False, 1)
# wgrad (A = X^T MN-major, B = dY MN-major), split-K
tot += bench("wgrad qkv", d, 3 * d, Me, True, True, 8, split=True)
tot += bench("wgrad d->d", d, d, Me, True, True, 17, split=True)
tot += bench("wgrad kv2", d, 2 * d, Me, True, True, 4, split=True)
tot += bench("wgrad ffn1", d, f, Me, True, True, 8, split=True)
tot += bench("wgrad ffn2", f, d, Me, True, True, 8, split=True)
tot += bench("wgrad logits", d, V, Md, True, True, 1, split=True)
print(f"sum over the step: {tot / 1e3:.2f} ms")

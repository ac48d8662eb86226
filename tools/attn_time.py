"""Times the attention kernels at cfg2 shapes (B=128, H=8, T=200/199, dh=32) through the C ABI: backend 2 = key-block
streaming kernels (skf_attn_tc2.cu), 6 = warp-specialised whole-row kernels (skf_attn_tc3.cu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from sketchformer_b200 import _lib
from sketchformer_b200.ops import ptr, stream

lib = _lib.load()
B, H, dh = 128, 8, 32
d = H * dh
backends = [int(x) for x in sys.argv[1:]] or [2, 6, 7]        # 7 = whole-row kernels with operand images
for (T, mode) in ((200, 1), (199, 2)):
    g = torch.Generator(device="cuda").manual_seed(0)
    qkv = torch.randn(B * T, 3 * d, device="cuda", generator=g)
    O = torch.empty(B * T, d, device="cuda"); dO = torch.randn(B * T, d, device="cuda", generator=g) * 1e-3
    lse = torch.empty(B, H, T, device="cuda"); delta = torch.empty(B, H, T, device="cuda")
    dqkv = torch.empty_like(qkv)
    pad = (torch.arange(T, device="cuda")[None, :] >= torch.randint(T // 3, T + 1, (B, 1), device="cuda")).to(torch.uint8).contiguous()

    img = torch.empty(lib.skf_attn_image_bytes(B, H, T, T, 0), dtype=torch.uint8, device="cuda")
    img_do = torch.empty(lib.skf_attn_image_bytes(B, H, T, T, 1), dtype=torch.uint8, device="cuda")

    def fwd(be):
        if be == 7:
            return _lib.check(lib.skf_attn_fwd_img(ptr(qkv), 3 * d, ptr(qkv[:, d:]), 3 * d, ptr(qkv[:, 2 * d:]), 3 * d, ptr(O), d,
                                                   ptr(lse), ptr(pad), B, H, T, T, dh, mode, ptr(img), stream()))
        _lib.check(lib.skf_attn_fwd(ptr(qkv), 3 * d, ptr(qkv[:, d:]), 3 * d, ptr(qkv[:, 2 * d:]), 3 * d, ptr(O), d, ptr(lse),
                                    ptr(pad), B, H, T, T, dh, mode, 0, be, stream()))

    def bwd(be):
        if be == 7:
            return _lib.check(lib.skf_attn_bwd_img(ptr(qkv), 3 * d, ptr(qkv[:, d:]), 3 * d, ptr(qkv[:, 2 * d:]), 3 * d, ptr(O), d,
                                                   ptr(dO), d, ptr(lse), ptr(delta), ptr(dqkv), 3 * d, ptr(dqkv[:, d:]), 3 * d,
                                                   ptr(dqkv[:, 2 * d:]), 3 * d, ptr(pad), B, H, T, T, dh, mode, ptr(img),
                                                   ptr(img_do), stream()))
        _lib.check(lib.skf_attn_bwd(ptr(qkv), 3 * d, ptr(qkv[:, d:]), 3 * d, ptr(qkv[:, 2 * d:]), 3 * d, ptr(O), d, ptr(dO), d,
                                    ptr(lse), ptr(delta), ptr(dqkv), 3 * d, ptr(dqkv[:, d:]), 3 * d, ptr(dqkv[:, 2 * d:]), 3 * d,
                                    ptr(pad), B, H, T, T, dh, mode, 0, be, stream()))

    for be in backends:
        for name, fn in (("fwd", fwd), ("bwd", bwd)):
            try:
                for _ in range(3):
                    fn(be)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(20):
                    fn(be)
                e1.record()
                torch.cuda.synchronize()
                print(f"T={T} mode={mode} backend {be} {name}: {e0.elapsed_time(e1) * 1e3 / 20:.1f} us")
            except Exception as e:      # noqa: BLE001
                print(f"T={T} mode={mode} backend {be} {name}: {str(e)[:120]}")

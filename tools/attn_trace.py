"""Phase trace of the whole-row attention forward kernel (CTA 0, first 8 items): prints clock deltas per role."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sketchformer_b200 import _lib  # noqa: E402
from sketchformer_b200.ops import ptr, stream  # noqa: E402

lib = _lib.load()
B, H, dh, T, mode = 128, 8, 32, int(sys.argv[1]) if len(sys.argv) > 1 else 200, int(sys.argv[2]) if len(sys.argv) > 2 else 1
d = H * dh
qkv = torch.randn(B * T, 3 * d, device="cuda")
O = torch.empty(B * T, d, device="cuda"); lse = torch.empty(B, H, T, device="cuda")
pad = (torch.arange(T, device="cuda")[None, :] >= torch.randint(T // 3, T + 1, (B, 1), device="cuda")).to(torch.uint8).contiguous()
which = sys.argv[3] if len(sys.argv) > 3 else "fwd"
dO = torch.randn(B * T, d, device="cuda") * 1e-3
delta = torch.empty(B, H, T, device="cuda"); dqkv = torch.empty_like(qkv)
img = torch.empty(lib.skf_attn_image_bytes(B, H, T, T, 0), dtype=torch.uint8, device="cuda")
img_do = torch.empty(lib.skf_attn_image_bytes(B, H, T, T, 1), dtype=torch.uint8, device="cuda")
for i in range(3):
    if i == 2:
        _lib.check(lib.skf_set_option(b"attn3_trace", 1 if which == "fwd" else (2 if which == "dq" else 3)))
    _lib.check(lib.skf_attn_fwd_img(ptr(qkv), 3 * d, ptr(qkv[:, d:]), 3 * d, ptr(qkv[:, 2 * d:]), 3 * d, ptr(O), d, ptr(lse),
                                    ptr(pad), B, H, T, T, dh, mode, ptr(img), stream()))
    if which != "fwd":
        _lib.check(lib.skf_attn_bwd_img(ptr(qkv), 3 * d, ptr(qkv[:, d:]), 3 * d, ptr(qkv[:, 2 * d:]), 3 * d, ptr(O), d, ptr(dO), d,
                                        ptr(lse), ptr(delta), ptr(dqkv), 3 * d, ptr(dqkv[:, d:]), 3 * d, ptr(dqkv[:, 2 * d:]), 3 * d,
                                        ptr(pad), B, H, T, T, dh, mode, ptr(img), ptr(img_do), stream()))
torch.cuda.synchronize()
print("key lengths of CTA 0's items:", [(int((pad[(i * 148) // H] == 0).sum())) for i in range(7)])
buf = (C.c_int64 * 512)()
n = lib.skf_debug_trace(buf, 512)
v = [buf[i] for i in range(n)]
t0 = min(x for x in v if x > 0)
if which != "fwd":
    names_b = {0: "MMA  per tile: [dS0 ready, acc free, acc0+next scores0 issued, dS1 ready, -, acc1+next scores1 issued]",
               1: "WG0  per tile: [S,dP ready, dS handed, acc ready, -]", 2: "WG1  per tile: [S,dP ready, dS handed, acc ready, -]",
               3: "CONV [buffers free, K/Q.., .., .., dO done | 5: before SB wait, 6: after SB wait]"}
names = names_b if which != "fwd" else {0: "MMA  [V ready | t0: P0a, Ofree, PVa issued, P0b, PVb+commit, next S issued | t1: ...]",
         1: "WG-A [S ready, pass1 done, P handed, O ready, epilogue done]",
         2: "WG-B [S ready, pass1 done, P handed, O ready, epilogue done]",
         3: "CONV [buffer free, Q done, K done, V done]"}
for role in range(4):
    print(names[role])
    for li in range(7):
        row = v[(role * 8 + li) * 16:(role * 8 + li) * 16 + 16]
        print(f"  item {li}: " + " ".join(f"{x - t0:7d}" if x > 0 else "      -" for x in row[:13]))

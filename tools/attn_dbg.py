"""Per-sample errors of the whole-row attention kernels under hand-made padding patterns (debug aid)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sketchformer_oracle as O  # noqa: E402
from sketchformer_b200 import _lib  # noqa: E402
from sketchformer_b200.ops import ptr, stream  # noqa: E402

lib = _lib.load()
B, H, Tq, Tk, dh = 6, 2, 200, 200, 32
d = H * dh
g = torch.Generator().manual_seed(5)
q = torch.randn(B, Tq, d, generator=g).double().requires_grad_(True)
k = torch.randn(B, Tk, d, generator=g).double().requires_grad_(True)
v = torch.randn(B, Tk, d, generator=g).double().requires_grad_(True)
pad = torch.zeros(B, Tk, dtype=torch.bool)
pad[0, 3:40] = True
pad[1, :] = True; pad[1, -1] = False
pad[2, :] = True; pad[2, 0] = False
pad[3, 150:] = True
pad[4, 100:] = True
pad[5, 128:] = True
mask = pad[:, None, None, :].double()
o, w = O.scaled_dot_product_attention(O.split_heads(q, H), O.split_heads(k, H), O.split_heads(v, H), mask)
o = o.permute(0, 2, 1, 3).reshape(B, Tq, d)
do = torch.randn(B, Tq, d, generator=g)
o.backward(do.double())
dev = lambda t: t.detach().float().cuda().contiguous()
qd, kd, vd, dod = dev(q), dev(k), dev(v), dev(do)
od = torch.empty(B, Tq, d, device="cuda"); lse = torch.empty(B, H, Tq, device="cuda")
padd = pad.to(torch.uint8).cuda().contiguous()
for be in (2, 6):
    _lib.check(lib.skf_attn_fwd(ptr(qd), d, ptr(kd), d, ptr(vd), d, ptr(od), d, ptr(lse), ptr(padd), B, H, Tq, Tk, dh, 1, 0, be, stream()))
    dq = torch.empty_like(qd); dk = torch.empty_like(kd); dv = torch.empty_like(vd); delta = torch.empty(B, H, Tq, device="cuda")
    _lib.check(lib.skf_attn_bwd(ptr(qd), d, ptr(kd), d, ptr(vd), d, ptr(od), d, ptr(dod), d, ptr(lse), ptr(delta), ptr(dq), d,
                                ptr(dk), d, ptr(dv), d, ptr(padd), B, H, Tq, Tk, dh, 1, 0, be, stream()))
    torch.cuda.synchronize()
    for b in range(B):
        e = lambda a_, r_: float((a_[b].double().cpu() - r_[b]).norm() / max(float(r_[b].norm()), 1e-30))
        print(f"backend {be} sample {b}: O {e(od, o.detach()):.2e} dQ {e(dq, q.grad):.2e} dK {e(dk, k.grad):.2e} dV {e(dv, v.grad):.2e}")

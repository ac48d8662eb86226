"""Micro-benchmarks of two HBM-bound kernels at cfg2 size (CUDA events, 50 launches after 5 warm-ups; the operands
rotate over 4 buffer sets so that nothing stays in L2): LayerNorm backward for several grid sizes
(skf_set_option ln_bwd_bpsm) and the embedding backward on a ragged synthetic batch."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sketchformer_b200 import _lib, synth                       # noqa: E402
from sketchformer_b200.ops import ptr, stream                   # noqa: E402

lib = _lib.load()
torch.manual_seed(0)
rows, d = 25600, 256


def timed(fn, n=50):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


sets = [(torch.randn(rows, d, device="cuda"), torch.randn(rows, d, device="cuda"), torch.empty(rows, d, device="cuda"))
        for _ in range(4)]
mean = torch.zeros(rows, device="cuda"); rstd = torch.ones(rows, device="cuda")
gamma = torch.ones(d, device="cuda"); dg = torch.zeros(d, device="cuda"); db = torch.zeros(d, device="cuda")
for nrows, bpsm in ((2, 2), (2, 3), (2, 4), (4, 1), (4, 2), (4, 3)):
    _lib.check(lib.skf_set_option(b"ln_bwd_bpsm", bpsm))
    _lib.check(lib.skf_set_option(b"ln_bwd_rows", nrows))

    def f(i):
        dy, s, ds = sets[i % 4]
        _lib.check(lib.skf_layernorm_bwd(ptr(dy), ptr(s), ptr(mean), ptr(rstd), ptr(gamma), ptr(ds), ptr(dg), ptr(db),
                                         rows, d, 0, stream()))
    us = timed(f)
    print(f"ln_bwd rows/warp={nrows} bpsm={bpsm}: {us:.1f} us  {3 * rows * d * 4 / us / 1e6:.2f} TB/s", flush=True)
_lib.check(lib.skf_set_option(b"ln_bwd_bpsm", 2))
_lib.check(lib.skf_set_option(b"ln_bwd_rows", 2))

x, _ = synth.token_batch(128, 200, 345, 1234, vocab=1004)
tok = torch.from_numpy(x).cuda()
douts = [torch.randn(128 * 200, d, device="cuda") for _ in range(4)]
table = torch.zeros(1004, d, device="cuda")
us = timed(lambda i: _lib.check(lib.skf_embed_bwd(ptr(tok), 200, ptr(douts[i % 4]), ptr(table), 128, 200, d, 1004, stream())))
print(f"embed_bwd: {us:.1f} us")

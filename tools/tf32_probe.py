"""Does the TMA unit round fp32->tf32 when the tensor map is typed TFLOAT32?  Prints the relative
error and the signed bias of an fp32 GEMM through kind::tf32 under both tensor-map data types."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sketchformer_b200 import _lib                       # noqa: E402
from tests.gpu_util import gemm, rel_err, to_tf32        # noqa: E402

lib = _lib.load()
g = torch.Generator().manual_seed(0)
M, N, K = 1024, 512, 512
A = torch.rand(M, K, generator=g) + 0.5       # positive operands expose truncation bias
B = torch.rand(K, N, generator=g) + 0.5
want = A.double() @ B.double()
for mode in (0, 1):
    _lib.check(lib.skf_set_option(b"tma_tf32", mode))
    for name, a, b in (("raw fp32", A, B), ("pre-rounded", to_tf32(A), to_tf32(B))):
        Y = torch.zeros(M, N, device="cuda")
        gemm(a.cuda(), K, 1, b.cuda(), N, 1, Y, N, M, N, K, 0)
        torch.cuda.synchronize()
        y = Y.double().cpu()
        print(f"tma_tf32={mode} {name:12s} rel_err={rel_err(y, want):.3e} signed_bias={float(((y - want) / want).mean()):+.3e}")
_lib.check(lib.skf_set_option(b"tma_tf32", 0))

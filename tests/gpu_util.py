"""Helpers shared by the GPU parity tests (all calls go through the C ABI of libskf.so)."""
import ctypes as C

import torch

from sketchformer_b200 import _lib
from sketchformer_b200._lib import SkfGemmEpilogue


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b||_2 / ||b||_2 in float64 (the parity metric of SURVEY §7 'hard parts')."""
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    den = float(b.norm())
    num = float((a - b).norm())
    return num / den if den > 0 else num


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def gemm(A, sa_m, sa_k, B, sb_k, sb_n, Y, ldy, M, N, K, backend, bias=None, resid=None, mask=None, act=0,
         accumulate=0, round_tf32=0, split_k=0):
    lib = _lib.load()
    e = SkfGemmEpilogue(bias=ptr(bias), resid=ptr(resid), ld_resid=(resid.stride(0) if resid is not None else 0),
                        mask=ptr(mask), ld_mask=(mask.stride(0) if mask is not None else 0), act=act,
                        accumulate=accumulate, round_tf32=round_tf32, split_k=split_k)
    _lib.check(lib.skf_gemm(ptr(A), sa_m, sa_k, ptr(B), sb_k, sb_n, ptr(Y), ldy, M, N, K, C.byref(e), backend,
                            stream()), "skf_gemm")


def to_tf32(x: torch.Tensor) -> torch.Tensor:
    """Round-to-nearest (ties away, like cvt.rna) fp32 -> tf32, on any device."""
    i = x.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF
    return i.view(torch.float32)

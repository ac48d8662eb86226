"""Thin Python wrappers over the operator-level C ABI (include/skf.h) for torch CUDA tensors."""
import ctypes as C

import torch

from . import _lib
from ._lib import SkfGemmEpilogue


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def gemm(A, sa_m, sa_k, B, sb_k, sb_n, Y, ldy, M, N, K, backend=0, bias=None, resid=None, mask=None, act=0,
         accumulate=0, round_tf32=0, split_k=0, a_lo=None, b_lo=None, colsum=None, b_hi16=None, b_lo16=None):
    """Y[M,N] = epilogue(op(A) op(B)); see skf_gemm in include/skf.h (Dense fwd / dgrad / wgrad)."""
    lib = _lib.load()
    e = SkfGemmEpilogue(bias=ptr(bias), resid=ptr(resid), ld_resid=(resid.stride(0) if resid is not None else 0),
                        mask=ptr(mask), ld_mask=(mask.stride(0) if mask is not None else 0), act=act,
                        accumulate=accumulate, round_tf32=round_tf32, split_k=split_k, a_lo=ptr(a_lo), b_lo=ptr(b_lo), colsum=ptr(colsum),
                        b_hi16=ptr(b_hi16), b_lo16=ptr(b_lo16))
    _lib.check(lib.skf_gemm(ptr(A), sa_m, sa_k, ptr(B), sb_k, sb_n, ptr(Y), ldy, M, N, K, C.byref(e), backend,
                            stream()), "skf_gemm")


def split_lo(x):
    """lo = x - tf32(x): low half of the 3xTF32 operand split (skf_split_lo)."""
    lo = torch.empty_like(x)
    _lib.check(_lib.load().skf_split_lo(ptr(x), ptr(lo), x.numel(), stream()), "skf_split_lo")
    return lo


def split_bf16(x):
    """(hi, lo) = (bf16(x), bf16(x - hi)): operand images of the bf16x3 product (skf_split_bf16)."""
    hi = torch.empty_like(x, dtype=torch.bfloat16)
    lo = torch.empty_like(x, dtype=torch.bfloat16)
    _lib.check(_lib.load().skf_split_bf16(ptr(x), ptr(hi), ptr(lo), x.numel(), stream()), "skf_split_bf16")
    return hi, lo

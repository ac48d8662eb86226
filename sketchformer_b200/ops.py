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


def _images(*ws):
    """bf16 hi / lo images of weights (or None each when `ws` holds None)"""
    out = []
    for w in ws:
        out.extend(split_bf16(w) if w is not None else (None, None))
    return out


def ffn_fwd(x, W1, b1, W2, b2, resid=None, images=True):
    """skf_ffn_fwd: (h, y) = (relu(x W1 + b1), resid + h W2 + b2); builders/layers/transformer.py:194-198, 220-222."""
    M, d = x.shape
    f = W1.shape[1]
    h = torch.empty(M, f, device=x.device)
    y = torch.empty(M, d, device=x.device)
    im = _images(W1, W2) if images else [None] * 4
    _lib.check(_lib.load().skf_ffn_fwd(ptr(x), ptr(W1), ptr(b1), ptr(W2), ptr(b2), ptr(resid), ptr(h), ptr(y), M, d, f,
                                       *[ptr(t) for t in im], stream()), "skf_ffn_fwd")
    return h, y


def ffn_bwd(x, h, dy, W1, W2, dx_resid=None, images=True):
    """skf_ffn_bwd -> (dx, dW1, db1, dW2, db2) for the gradient dy of the sublayer output."""
    M, d = x.shape
    f = W1.shape[1]
    dh = torch.empty(M, f, device=x.device)
    dx = torch.empty(M, d, device=x.device)
    dW1, db1 = torch.zeros_like(W1), torch.zeros(f, device=x.device)
    dW2, db2 = torch.zeros_like(W2), torch.zeros(d, device=x.device)
    im = _images(W1, W2) if images else [None] * 4
    _lib.check(_lib.load().skf_ffn_bwd(ptr(x), ptr(h), ptr(dy), ptr(W1), ptr(W2), ptr(dh), ptr(dx), ptr(dx_resid), ptr(dW1),
                                       ptr(db1), ptr(dW2), ptr(db2), M, d, f, *[ptr(t) for t in im], stream()), "skf_ffn_bwd")
    return dx, dW1, db1, dW2, db2


def logits_ce_fwd(x, W, b, tgt, grad_scale, images=True):
    """skf_logits_ce_fwd -> (dlogits [B*Td, V], stats double[2] = (CE sum, correct), argmax int32 [B*Td]);
    models/sketchformer.py:179 + builders/losses.py:26-41."""
    B, Td = tgt.shape
    d, V = W.shape
    logits = torch.empty(B * Td, V, device=x.device)
    stats = torch.zeros(2, dtype=torch.float64, device=x.device)
    am = torch.empty(B * Td, dtype=torch.int32, device=x.device)
    im = _images(W) if images else [None] * 2
    _lib.check(_lib.load().skf_logits_ce_fwd(ptr(x), ptr(W), ptr(b), ptr(tgt), tgt.stride(0), B, Td, d, V, float(grad_scale),
                                             ptr(logits), V, ptr(stats), ptr(am), *[ptr(t) for t in im], stream()),
               "skf_logits_ce_fwd")
    return logits, stats, am


def logits_ce_bwd(x, dlogits, W, B, Td, images=True):
    """skf_logits_ce_bwd -> (dx, dW, db)."""
    d, V = W.shape
    dx = torch.empty(B * Td, d, device=x.device)
    dW, db = torch.zeros_like(W), torch.zeros(V, device=x.device)
    im = _images(W) if images else [None] * 2
    _lib.check(_lib.load().skf_logits_ce_bwd(ptr(x), ptr(dlogits), dlogits.stride(0), ptr(W), ptr(dx), ptr(dW), ptr(db), B, Td,
                                             d, V, *[ptr(t) for t in im], stream()), "skf_logits_ce_bwd")
    return dx, dW, db

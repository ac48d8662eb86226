"""Operator-level parity on a B200: every C-ABI kernel against the oracle's op (or its direct
torch-CPU fp64 equivalent) on seeded inputs.  Integer/index outputs must be bit-exact; fp32
kernels within 1e-5 (SIMT, exact fp32) or 1e-3 (tcgen05 TF32), the north-star's tolerance."""
import ctypes as C
import math

import pytest
import torch

from oracle import sketchformer_oracle as O
from sketchformer_b200 import _lib, synth
from sketchformer_b200.ops import split_bf16
from tests.gpu_util import gemm, ptr, rel_err, stream, to_tf32

pytestmark = pytest.mark.gpu

TOL_EXACT = 2e-5   # fp32 SIMT kernels vs fp64 oracle
TOL_TF32 = 1e-3    # tensor-core kernels (north-star: 1e-3 fp32 relative)


@pytest.fixture(scope="module")
def lib():
    lib = _lib.load()
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    assert lib.skf_device_ok() == 1, "not an sm_100 device"
    return lib


def dev(t):
    return t.cuda().contiguous()


def test_positional_encoding_and_embed(lib):
    B, T, d, V = 5, 37, 128, 1004
    pe = torch.empty(T, d, device="cuda")
    _lib.check(lib.skf_positional_encoding(ptr(pe), T, d, stream()))
    ref = O.positional_encoding(1000, d)[0, :T]
    # float64 sin/cos on device vs numpy: at most 1 ulp of fp32 after the final rounding
    assert float((pe.cpu() - ref).abs().max()) <= 1.2e-7
    x, _ = synth.token_batch(B, T, 10, 3)
    tok = dev(torch.from_numpy(x))
    g = torch.Generator().manual_seed(0)
    table = (torch.rand(V, d, generator=g) - 0.5) * 0.1
    out = torch.empty(B, T, d, device="cuda")
    table_d, ref_d = dev(table), dev(ref)
    _lib.check(lib.skf_embed_fwd(ptr(tok), T, ptr(table_d), ptr(ref_d), ptr(out), B, T, d, V, 0, stream()))
    p = {"enc.emb": table}
    cfg = O.OracleConfig(d_model=d, max_seq_len=T)
    want = O.embed_input(p, "enc", torch.from_numpy(x), cfg, torch.float32)
    assert torch.equal(out.cpu(), want), "embedding gather*sqrt(d)+PE must be bit-exact in fp32"
    # decoder view: tar_inp = tok[:, :-1] with the row stride of the full batch
    out2 = torch.empty(B, T - 1, d, device="cuda")
    _lib.check(lib.skf_embed_fwd(ptr(tok), T, ptr(table_d), ptr(ref_d), ptr(out2), B, T - 1, d, V, 0, stream()))
    assert torch.equal(out2.cpu(), want[:, :-1])
    # pad flags (builders/utils.py:35-43)
    pad = torch.empty(B, T, dtype=torch.uint8, device="cuda")
    _lib.check(lib.skf_pad_flags(ptr(tok), T, ptr(pad), B, T, stream()))
    assert torch.equal(pad.cpu().bool(), torch.from_numpy(x) == 0)
    # backward scatter-add (duplicates accumulate)
    dout = torch.randn(B, T, d, generator=g)
    dtab = torch.zeros(V, d, device="cuda")
    dout_d = dev(dout)
    _lib.check(lib.skf_embed_bwd(ptr(tok), T, ptr(dout_d), ptr(dtab), B, T, d, V, stream()))
    want_g = torch.zeros(V, d, dtype=torch.float64)
    want_g.index_add_(0, torch.from_numpy(x).reshape(-1), dout.double().reshape(-1, d) * math.sqrt(d))
    assert rel_err(dtab, want_g) < TOL_EXACT


@pytest.mark.parametrize("rows,d", [(77, 64), (300, 128), (1000, 256), (129, 512)])
def test_layernorm_fwd_bwd(lib, rows, d):
    g = torch.Generator().manual_seed(rows)
    a = torch.randn(rows, d, generator=g)
    b = torch.randn(rows, d, generator=g)
    bias = torch.randn(d, generator=g)
    gamma = torch.rand(d, generator=g) + 0.5
    beta = torch.randn(d, generator=g)
    s = torch.empty(rows, d, device="cuda"); y = torch.empty_like(s)
    mean = torch.empty(rows, device="cuda"); rstd = torch.empty_like(mean)
    a_d, b_d, bias_d, gamma_d, beta_d = dev(a), dev(b), dev(bias), dev(gamma), dev(beta)
    _lib.check(lib.skf_add_layernorm_fwd(ptr(a_d), ptr(b_d), ptr(bias_d), ptr(gamma_d), ptr(beta_d),
                                         ptr(s), ptr(y), ptr(mean), ptr(rstd), rows, d, 0, stream()))
    sd = (a.double() + (b.double() + bias.double())).requires_grad_(True)
    p = {"ln.g": gamma.double().requires_grad_(True), "ln.b": beta.double().requires_grad_(True)}
    want = O.layer_norm(p, "ln", sd)
    assert rel_err(s, sd) < 1e-6 and rel_err(y, want) < TOL_EXACT
    dy = torch.randn(rows, d, generator=g)
    want.backward(dy.double())
    ds = torch.empty(rows, d, device="cuda")
    dg = torch.zeros(d, device="cuda"); db = torch.zeros(d, device="cuda")
    dy_d = dev(dy)
    _lib.check(lib.skf_layernorm_bwd(ptr(dy_d), ptr(s), ptr(mean), ptr(rstd), ptr(gamma_d), ptr(ds), ptr(dg),
                                     ptr(db), rows, d, 0, stream()))
    assert rel_err(ds, sd.grad) < TOL_EXACT
    assert rel_err(dg, p["ln.g"].grad) < TOL_EXACT and rel_err(db, p["ln.b"].grad) < TOL_EXACT


def test_dropout_mask_matches_numpy_mirror(lib):
    """skf_dropout: counter-based keep mask == oracle/dropout_rng.py bit for bit; keep fraction ~ 1 - rate."""
    from oracle.dropout_rng import keep_mask
    n, rate, seed, site = 257 * 129, 0.1, 1234567, 102
    x = torch.randn(n)
    step = torch.tensor([4321], dtype=torch.int64, device="cuda")
    y = torch.empty(n, device="cuda")
    xd = dev(x)
    _lib.check(lib.skf_dropout(ptr(xd), ptr(y), n, rate, seed, site, ptr(step), stream()))
    keep = torch.from_numpy(keep_mask((n,), rate, seed, site, 4321))
    want = torch.where(keep, x * (torch.tensor(1.0) / (torch.tensor(1.0) - torch.tensor(rate))), torch.zeros(()))
    assert torch.equal(y.cpu(), want)
    assert abs(float(keep.float().mean()) - 0.9) < 0.01
    step += 1
    y2 = torch.empty(n, device="cuda")
    _lib.check(lib.skf_dropout(ptr(xd), ptr(y2), n, rate, seed, site, ptr(step), stream()))
    assert not torch.equal(y2, y)


def test_layernorm_dropout_fwd_bwd(lib):
    """s = a + dropout(b + bias); y = LN(s); backward returns ds (residual path) and mask(ds) (sublayer path)."""
    from oracle.dropout_rng import keep_mask
    rows, d, rate, seed, site = 300, 256, 0.25, 99, 11
    g = torch.Generator().manual_seed(5)
    a, b = torch.randn(rows, d, generator=g), torch.randn(rows, d, generator=g)
    bias, gamma, beta = torch.randn(d, generator=g), torch.randn(d, generator=g), torch.randn(d, generator=g)
    dy = torch.randn(rows, d, generator=g)
    step = torch.tensor([7], dtype=torch.int64, device="cuda")
    keep = torch.from_numpy(keep_mask((rows, d), rate, seed, site, 7)).double()
    a64 = a.double().requires_grad_(True); b64 = b.double().requires_grad_(True)
    s64 = a64 + (b64 + bias.double()) * keep / (1 - rate)
    y64 = torch.nn.functional.layer_norm(s64, (d,), gamma.double(), beta.double(), 1e-6)
    y64.backward(dy.double())
    s, y = torch.empty(rows, d, device="cuda"), torch.empty(rows, d, device="cuda")
    mean, rstd = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    ad, bd, biasd, gd, bed, dyd = dev(a), dev(b), dev(bias), dev(gamma), dev(beta), dev(dy)   # keep the buffers alive
    _lib.check(lib.skf_add_dropout_layernorm_fwd(ptr(ad), ptr(bd), ptr(biasd), ptr(gd), ptr(bed),
                                                 ptr(s), ptr(y), ptr(mean), ptr(rstd), rows, d, rate, seed, site,
                                                 ptr(step), stream()))
    assert rel_err(s, s64) < 1e-6 and rel_err(y, y64) < 1e-5
    ds, dsd = torch.empty(rows, d, device="cuda"), torch.empty(rows, d, device="cuda")
    dg, db = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
    _lib.check(lib.skf_layernorm_dropout_bwd(ptr(dyd), ptr(s), ptr(mean), ptr(rstd), ptr(gd), ptr(ds),
                                             ptr(dsd), ptr(dg), ptr(db), rows, d, rate, seed, site, ptr(step), stream()))
    assert rel_err(ds, a64.grad) < 1e-5 and rel_err(dsd, b64.grad) < 1e-5


def test_colsum(lib):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1234, 345, generator=g)
    out = torch.zeros(345, device="cuda")
    xd = dev(x)
    _lib.check(lib.skf_colsum(ptr(xd), 345, ptr(out), 1234, 345, stream()))
    assert rel_err(out, x.double().sum(0)) < TOL_EXACT


@pytest.mark.parametrize("B,T,d,U", [(6, 50, 128, 64), (3, 5, 32, 16), (5, 200, 256, 256), (2, 64, 128, 128),
                                     (2, 33, 2048, 64)])
def test_pool_fwd_bwd(lib, B, T, d, U):
    """SelfAttnV1 pooling (transformer.py:61-74) on clusters of 8 CTAs per sample (T not a multiple of 8, T < 8:
    CTAs without rows) and, when the cluster's gather buffer would not fit (d = 2048), on one CTA per sample."""
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, T, d, generator=g).double().requires_grad_(True)
    p = {"pool.W": (torch.randn(d, U, generator=g) * 0.05).double().requires_grad_(True),
         "pool.b": (torch.randn(U, generator=g) * 0.1).double().requires_grad_(True),
         "pool.V": ((torch.rand(U, 1, generator=g) - 0.5)).double().requires_grad_(True)}
    cfg = O.OracleConfig(d_model=d, lowerdim=U, max_seq_len=T)
    emb, ai = O.bottleneck(p, x, cfg)
    u = torch.tanh(x.detach() @ p["pool.W"].detach() + p["pool.b"].detach()).float()
    xd, ud, vd = dev(x.detach().float()), dev(u), dev(p["pool.V"].detach().float().reshape(-1))
    a = torch.empty(B, T, device="cuda"); e = torch.empty(B, d, device="cuda")
    _lib.check(lib.skf_pool_fwd(ptr(xd), ptr(ud), ptr(vd), ptr(a), ptr(e), B, T, d, U, 0, stream()))
    assert rel_err(a, ai[..., 0]) < TOL_EXACT and rel_err(e, emb) < TOL_EXACT
    demb = torch.randn(B, d, generator=g)
    emb.backward(demb.double())
    dx = torch.empty(B, T, d, device="cuda"); dupre = torch.empty(B, T, U, device="cuda")
    dv = torch.zeros(U, device="cuda")
    demb_d = dev(demb)
    _lib.check(lib.skf_pool_bwd(ptr(xd), ptr(ud), ptr(vd), ptr(a), ptr(demb_d), ptr(dx), 0, ptr(dupre), ptr(dv),
                                B, T, d, U, 0, stream()))
    assert rel_err(dv, p["pool.V"].grad.reshape(-1)) < TOL_EXACT
    # dW = x^T dupre, db = colsum(dupre), dx_total = dx + dupre W^T
    dup = dupre.double().cpu()
    assert rel_err(x.detach().reshape(-1, d).T @ dup.reshape(-1, U), p["pool.W"].grad) < TOL_EXACT
    assert rel_err(dup.reshape(-1, U).sum(0), p["pool.b"].grad) < TOL_EXACT
    assert rel_err(dx.double().cpu() + dup @ p["pool.W"].detach().T, x.grad) < TOL_EXACT


def test_expand_fwd_bwd(lib):
    B, T, d = 5, 33, 64
    g = torch.Generator().manual_seed(3)
    emb = torch.randn(B, d, generator=g).double().requires_grad_(True)
    p = {"exp.k": torch.randn(1, T, generator=g).double().requires_grad_(True),
         "exp.b": torch.randn(T, generator=g).double().requires_grad_(True)}
    pre = O.dense_expander(p, emb)
    out = torch.empty(B, T, d, device="cuda")
    ed, wd, bd = dev(emb.detach().float()), dev(p["exp.k"].detach().float().reshape(-1)), dev(p["exp.b"].detach().float())
    _lib.check(lib.skf_expand_fwd(ptr(ed), ptr(wd), ptr(bd), ptr(out), B, T, d, 0, stream()))
    assert rel_err(out, pre) < 1e-6
    dpre = torch.randn(B, T, d, generator=g)
    pre.backward(dpre.double())
    demb = torch.empty(B, d, device="cuda"); dw = torch.zeros(T, device="cuda"); db = torch.zeros(T, device="cuda")
    dpre_d = dev(dpre)
    _lib.check(lib.skf_expand_bwd(ptr(dpre_d), ptr(ed), ptr(wd), ptr(demb), 0, ptr(dw), ptr(db), B, T, d, stream()))
    assert rel_err(demb, emb.grad) < TOL_EXACT
    assert rel_err(dw, p["exp.k"].grad.reshape(-1)) < TOL_EXACT and rel_err(db, p["exp.b"].grad) < TOL_EXACT


def test_token_and_class_ce(lib):
    B, Td, V = 7, 29, 1004
    x, y = synth.token_batch(B, Td + 1, 345, 11)
    g = torch.Generator().manual_seed(4)
    logits = (torch.randn(B, Td, V, generator=g) * 2).double().requires_grad_(True)
    tgt = torch.from_numpy(x)[:, 1:]
    loss = O.reconstruction_loss(tgt, logits)
    loss.backward()
    tok = dev(torch.from_numpy(x))
    lg = dev(logits.detach().float())
    dl = torch.empty_like(lg)
    stats = torch.zeros(2, dtype=torch.float64, device="cuda")
    am = torch.empty(B * Td, dtype=torch.int32, device="cuda")
    scale = 1.0 / (B * Td)
    _lib.check(lib.skf_token_ce(ptr(lg), V, C.c_void_p(tok.data_ptr() + 8), Td + 1, B, Td, V, scale, ptr(dl), V,
                                ptr(stats), ptr(am), 0, stream()))
    st = stats.cpu()
    assert abs(float(st[0]) * scale - float(loss)) < 1e-5 * abs(float(loss))
    want_am = logits.detach().float().argmax(-1).reshape(-1).to(torch.int32)
    assert torch.equal(am.cpu(), want_am), "argmax must be bit-exact"
    assert float(st[1]) == float((want_am.long() == tgt.reshape(-1)).sum())
    assert rel_err(dl, logits.grad) < TOL_EXACT
    # classification head
    Cn = 345
    cl = (torch.randn(B, Cn, generator=g)).double().requires_grad_(True)
    lab = torch.from_numpy(y)
    closs = O.class_loss(lab, cl)
    closs.backward()
    Cp = 348
    clp = torch.zeros(B, Cp); clp[:, :Cn] = cl.detach().float()
    cld = dev(clp)
    dcl = torch.zeros(B, Cp, device="cuda"); probs = torch.empty(B, Cn, device="cuda")
    cst = torch.zeros(2, dtype=torch.float64, device="cuda"); cam = torch.empty(B, dtype=torch.int32, device="cuda")
    lab_d = dev(lab)
    _lib.check(lib.skf_class_ce(ptr(cld), Cp, ptr(lab_d), B, Cn, 1.0 / B, ptr(dcl), Cp, ptr(cst), ptr(cam),
                                ptr(probs), stream()))
    assert abs(float(cst.cpu()[0]) / B - float(closs)) < 1e-5 * abs(float(closs))
    assert rel_err(dcl[:, :Cn], cl.grad) < TOL_EXACT
    assert rel_err(probs, torch.softmax(cl.detach(), -1)) < TOL_EXACT
    assert torch.equal(cam.cpu().long(), cl.detach().argmax(-1))


def test_adam_matches_oracle(lib):
    n = 10007
    g = torch.Generator().manual_seed(5)
    w = {"a": torch.randn(n, generator=g)}
    opt = O.AdamTF({"a": w["a"].double()}, 128)
    w64 = {"a": w["a"].double().clone()}
    pd = dev(w["a"]); m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda")
    step = torch.zeros(1, dtype=torch.int64, device="cuda")
    for it in range(5):
        grad = torch.randn(n, generator=g) * 0.01
        grad_d = dev(grad)
        _lib.check(lib.skf_adam_step(ptr(pd), ptr(grad_d), ptr(m), ptr(v), n, ptr(step), 128, 5000, 0.9, 0.98, 1e-9,
                                     1.0, stream()))
        opt.apply(w64, {"a": grad.double()})
        if it == 0:   # lr(0) == 0: weights bit-identical, moments moved
            assert torch.equal(pd.cpu(), w["a"]) and float(m.abs().sum()) > 0
            assert int(step.item()) == 1
            # jump into the warm-up ramp (lr ~ 1e-3) so that the displacement is resolvable in fp32
            step.fill_(4000); opt.iterations = 4000
    assert int(step.item()) == 4004
    # the update magnitude is ~lr*1; compare the accumulated displacement
    assert rel_err(pd.cpu().double() - w["a"].double(), w64["a"] - w["a"].double()) < 1e-4
    assert rel_err(m, opt.m["a"]) < 1e-5 and rel_err(v, opt.v["a"]) < 1e-5


def _attn_case(lib, B, H, Tq, Tk, dh, mode, seed, fwd_backend=1, bwd_backend=1, pad_edit=None, len_range=None):
    g = torch.Generator().manual_seed(seed)
    d = H * dh
    q = torch.randn(B, Tq, d, generator=g).double().requires_grad_(True)
    k = torch.randn(B, Tk, d, generator=g).double().requires_grad_(True)
    v = torch.randn(B, Tk, d, generator=g).double().requires_grad_(True)
    lo, hi = len_range or (max(1, Tk // 3), Tk)
    lens = torch.randint(lo, hi + 1, (B,), generator=g)
    pad = (torch.arange(Tk)[None, :] >= lens[:, None])
    if pad_edit is not None:
        pad = pad_edit(pad.clone())
    mask = None
    if mode >= 1:
        mask = pad[:, None, None, :].double()
    if mode == 2:
        mask = torch.maximum(mask, O.create_look_ahead_mask(Tq, torch.float64))
    o, w = O.scaled_dot_product_attention(O.split_heads(q, H), O.split_heads(k, H), O.split_heads(v, H), mask)
    o = o.permute(0, 2, 1, 3).reshape(B, Tq, d)
    do = torch.randn(B, Tq, d, generator=g)
    o.backward(do.double())
    qd, kd, vd = dev(q.detach().float()), dev(k.detach().float()), dev(v.detach().float())
    od = torch.empty(B, Tq, d, device="cuda"); lse = torch.empty(B, H, Tq, device="cuda")
    padd = dev(pad.to(torch.uint8)) if mode >= 1 else None
    img = img_do = None
    if fwd_backend == 7:          # whole-row kernels with operand images (what the plan runs in training)
        img = torch.empty(lib.skf_attn_image_bytes(B, H, Tq, Tk, 0), dtype=torch.uint8, device="cuda")
        img_do = torch.empty(lib.skf_attn_image_bytes(B, H, Tq, Tk, 1), dtype=torch.uint8, device="cuda")
        _lib.check(lib.skf_attn_fwd_img(ptr(qd), d, ptr(kd), d, ptr(vd), d, ptr(od), d, ptr(lse), ptr(padd), B, H, Tq, Tk,
                                        dh, mode, ptr(img), stream()))
    else:
        _lib.check(lib.skf_attn_fwd(ptr(qd), d, ptr(kd), d, ptr(vd), d, ptr(od), d, ptr(lse), ptr(padd), B, H, Tq, Tk, dh,
                                    mode, 0, fwd_backend, stream()))
    e_o = rel_err(od, o)
    e_lse = rel_err(lse, torch.logsumexp((O.split_heads(q, H) @ O.split_heads(k, H).transpose(-1, -2)).detach() / dh ** 0.5
                                         + (mask * -1e9 if mask is not None else 0), -1))
    e_o = max(e_o, e_lse)
    dq = torch.empty_like(qd); dk = torch.empty_like(kd); dv = torch.empty_like(vd)
    delta = torch.empty(B, H, Tq, device="cuda")
    do_d = dev(do)
    if fwd_backend == 7:
        _lib.check(lib.skf_attn_bwd_img(ptr(qd), d, ptr(kd), d, ptr(vd), d, ptr(od), d, ptr(do_d), d, ptr(lse), ptr(delta),
                                        ptr(dq), d, ptr(dk), d, ptr(dv), d, ptr(padd), B, H, Tq, Tk, dh, mode, ptr(img),
                                        ptr(img_do), stream()))
    else:
        _lib.check(lib.skf_attn_bwd(ptr(qd), d, ptr(kd), d, ptr(vd), d, ptr(od), d, ptr(do_d), d, ptr(lse), ptr(delta),
                                    ptr(dq), d, ptr(dk), d, ptr(dv), d, ptr(padd), B, H, Tq, Tk, dh, mode, 0, bwd_backend, stream()))
    return e_o, rel_err(dq, q.grad), rel_err(dk, k.grad), rel_err(dv, v.grad)


@pytest.mark.parametrize("B,H,Tq,Tk,dh,mode", [(2, 4, 24, 24, 16, 1), (2, 4, 23, 23, 16, 2), (2, 4, 23, 24, 16, 0),
                                               (3, 8, 200, 200, 32, 1), (3, 8, 199, 199, 32, 2),
                                               (3, 8, 199, 200, 32, 0), (2, 8, 130, 131, 64, 2),
                                               (2, 2, 70, 257, 64, 1)])
def test_attention_simt(lib, B, H, Tq, Tk, dh, mode):
    if mode == 2 and Tq != Tk:
        Tk = Tq
    errs = _attn_case(lib, B, H, Tq, Tk, dh, mode, 100 + Tq)
    assert max(errs) < TOL_EXACT, errs


TC_ATTN = [(3, 8, 200, 200, 32, 1), (3, 8, 199, 199, 32, 2), (3, 8, 199, 200, 32, 0), (2, 4, 64, 64, 32, 1),
           (2, 2, 130, 100, 32, 0), (2, 4, 23, 24, 32, 0), (1, 2, 224, 224, 32, 2)]


# 3xTF32 score products; P / dS rounded once to TF32 (rel. 2^-12 per weight, does not average out against
# random V) with compensated V / K / Q / dO: 1.6e-4 forward, 2.2e-4 backward measured on B200
TOL_ATTN_X3 = 3e-4
ATTN3_BWD = 6       # backward kernels the whole-row forward is paired with in the op test (6 once they exist)


# head dim 64 (two 128-byte chunks per tile) and sequences up to 1024 (cfg5: d 512, 8 heads, T 512)
TC_ATTN_WIDE = [(2, 8, 130, 131, 64, 2), (2, 2, 70, 257, 64, 1), (2, 3, 200, 200, 64, 0), (1, 8, 512, 512, 64, 2),
                (1, 2, 512, 512, 64, 1), (1, 2, 1000, 1000, 32, 2), (1, 1, 300, 1024, 64, 1), (2, 2, 1, 65, 64, 1)]


@pytest.mark.parametrize("B,H,Tq,Tk,dh,mode", TC_ATTN + TC_ATTN_WIDE +
                         [(2, 8, 256, 256, 32, 2), (2, 2, 1, 200, 32, 1), (3, 4, 65, 64, 32, 0)])
def test_attention_tcgen05(lib, B, H, Tq, Tk, dh, mode):
    """flash-style tcgen05 kernels: fp32-grade (backend 2) and single-pass TF32 (backend 0)."""
    if mode == 2 and Tq != Tk:
        Tk = Tq
    errs = _attn_case(lib, B, H, Tq, Tk, dh, mode, 200 + Tq, fwd_backend=2, bwd_backend=2)
    assert max(errs) < TOL_ATTN_X3, errs
    errs = _attn_case(lib, B, H, Tq, Tk, dh, mode, 300 + Tq, fwd_backend=0, bwd_backend=0)
    assert max(errs) < TOL_TF32, errs


# warp-specialised whole-row kernels (skf_attn_tc3.cu, backend 6): head dim 32, sequences up to 224; every product is
# the bf16x3 error-compensated sum (16 mantissa bits per operand): 1e-5-grade
TOL_ATTN3 = 5e-5
TC_ATTN3 = TC_ATTN + [(5, 8, 200, 200, 32, 1), (40, 8, 199, 199, 32, 2), (2, 3, 128, 128, 32, 2), (2, 3, 129, 129, 32, 2),
                      (1, 1, 1, 1, 32, 0), (2, 2, 1, 200, 32, 1), (2, 2, 224, 17, 32, 1), (3, 4, 65, 64, 32, 0)]


@pytest.mark.parametrize("B,H,Tq,Tk,dh,mode", TC_ATTN3)
def test_attention_tcgen05_whole_row(lib, B, H, Tq, Tk, dh, mode):
    if mode == 2 and Tq != Tk:
        Tk = Tq
    e_o, e_q, e_k, e_v = _attn_case(lib, B, H, Tq, Tk, dh, mode, 400 + Tq, fwd_backend=6, bwd_backend=ATTN3_BWD)
    assert e_o < TOL_ATTN3, e_o
    assert max(e_q, e_k, e_v) < (TOL_ATTN3 if ATTN3_BWD == 6 else TOL_ATTN_X3), (e_q, e_k, e_v)


@pytest.mark.parametrize("B,H,Tq,Tk,dh,mode", [c for c in TC_ATTN3 if c[2] == c[3] or c[5] == 2])
def test_attention_tcgen05_whole_row_with_images(lib, B, H, Tq, Tk, dh, mode):
    """The training configuration of the whole-row kernels: the forward pass writes the bf16 operand images, the
    backward kernels load them directly (skf_attn_fwd_img / skf_attn_bwd_img)."""
    if mode == 2 and Tq != Tk:
        Tk = Tq
    errs = _attn_case(lib, B, H, Tq, Tk, dh, mode, 500 + Tq, fwd_backend=7, bwd_backend=7)
    assert max(errs) < TOL_ATTN3, errs


def test_attention_whole_row_padding_patterns(lib):
    """The whole-row kernels drop the key columns (dK / dV: key tiles) past an item's last unpadded key.  Masks that
    are not a padded tail must still follow the reference's additive -1e9 exactly (builders/utils.py:93-97): padded
    keys in the MIDDLE of a sequence, real keys only at the very end, very short effective lengths, and lengths at the
    tile / half boundaries of the kernels."""
    def edit(pad):
        pad[:] = False
        pad[0, 3:40] = True                 # hole before the tail
        pad[1, :-2] = True                  # only the last two keys are real
        pad[2, 2:] = True                   # effective length 2
        pad[3, 128:] = True                 # effective length 128: key tile 1 skipped exactly at the tile edge
        pad[4, 100:] = True                 # effective length 100: key tile 1 skipped, key half 1 of dq empty
        pad[5, 113:] = True; pad[5, 7] = True
        return pad
    # (Not here: a single real key makes dQ = dK = 0 identically, and with EVERY key padded the fp32 reference is
    # uniform over the keys — 1e9 swallows the logits — while the fp64 oracle still sees them: neither has a
    # meaningful relative error.  DESIGN.md section 3.)
    for Tq, Tk in ((200, 200), (70, 200), (200, 131)):
        e_o, e_q, e_k, e_v = _attn_case(lib, 6, 2, Tq, Tk, 32, 1, 900 + Tq + Tk, fwd_backend=6, bwd_backend=6, pad_edit=edit)
        assert max(e_o, e_q, e_k, e_v) < TOL_ATTN3, (Tq, Tk, e_o, e_q, e_k, e_v)
    e_o, e_q, e_k, e_v = _attn_case(lib, 6, 2, 200, 200, 32, 1, 1300, fwd_backend=7, bwd_backend=7, pad_edit=edit)
    assert max(e_o, e_q, e_k, e_v) < TOL_ATTN3, ("images", e_o, e_q, e_k, e_v)


@pytest.mark.parametrize("backend", [6, 7])
@pytest.mark.parametrize("T,mode", [(200, 1), (199, 2)])
@pytest.mark.parametrize("len_range", [(2, 40), (23, 140), (100, 135)])
def test_attention_whole_row_short_sequences_many_items(lib, backend, T, mode, len_range):
    """Several (sample, head) items per CTA with SHORT sequences (the lengths the stroke3-distributed loader yields;
    the synthetic batches never go below T / 4): a unit's row math is then shorter than the MMA issue, the regime in
    which the dq kernel's warpgroups once read each other's accumulator phases."""
    errs = _attn_case(lib, 64, 8, T, T, 32, mode, 1700 + T + len_range[0], fwd_backend=backend, bwd_backend=backend,
                      len_range=len_range)
    assert max(errs) < TOL_ATTN3, errs


MAJORS = [("fwd", False, True), ("dgrad", False, False), ("wgrad", True, True), ("tn", True, False)]


def _gemm_case(M, N, K, a_mn, b_mn, backend, seed, **epi):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(M, K, generator=g)
    Bm = torch.randn(K, N, generator=g)
    if backend == 0:
        A, Bm = to_tf32(A), to_tf32(Bm)      # exactly TF32-representable operands: only accumulation order differs
    Ast = dev(A.T) if a_mn else dev(A)        # MN-major: stored [K, M]
    Bst = dev(Bm) if b_mn else dev(Bm.T)      # K-major: stored [N, K]
    sa = (1, M) if a_mn else (K, 1)
    sb = (N, 1) if b_mn else (1, K)
    Y = torch.zeros(M, N, device="cuda")
    kw = {}
    want = A.double() @ Bm.double()
    if epi.get("bias"):
        bias = torch.randn(N, generator=g); kw["bias"] = dev(bias); want = want + bias.double()
    if epi.get("act") == 1:
        kw["act"] = 1; want = torch.relu(want)
    if epi.get("act") == 2:
        kw["act"] = 2; want = torch.tanh(want)
    if epi.get("mask"):
        mk = torch.randn(M, N, generator=g); kw["mask"] = dev(mk); want = want * (mk > 0)
    if epi.get("resid"):
        rs = torch.randn(M, N, generator=g); kw["resid"] = dev(rs); want = want + rs.double()
    if epi.get("split_k"):
        kw["accumulate"] = 1; kw["split_k"] = epi["split_k"]
    if epi.get("images"):        # B is a weight with bf16 images in its own layout (skf_split_bf16)
        kw["b_hi16"], kw["b_lo16"] = split_bf16(Bst)
    gemm(Ast, sa[0], sa[1], Bst, sb[0], sb[1], Y, N, M, N, K, backend, **kw)
    torch.cuda.synchronize()
    return rel_err(Y, want)


@pytest.mark.parametrize("name,a_mn,b_mn", MAJORS)
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (200, 100, 96), (333, 768, 256), (64, 1004, 128)])
def test_gemm_simt(lib, name, a_mn, b_mn, M, N, K):
    assert _gemm_case(M, N, K, a_mn, b_mn, 1, 7) < TOL_EXACT
    assert _gemm_case(M, N, K, a_mn, b_mn, 1, 8, bias=True, act=1, resid=True) < TOL_EXACT
    assert _gemm_case(M, N, K, a_mn, b_mn, 1, 9, mask=True) < TOL_EXACT
    assert _gemm_case(M, N, K, a_mn, b_mn, 1, 10, split_k=3) < TOL_EXACT


TC_SHAPES = [(128, 256, 32), (128, 64, 256), (256, 128, 64), (200, 100, 96), (1600, 768, 256), (512, 1004, 128),
             (256, 256, 4000), (3000, 1024, 256)]


@pytest.mark.parametrize("backend", [0], ids=["persistent"])
@pytest.mark.parametrize("name,a_mn,b_mn", MAJORS)
@pytest.mark.parametrize("M,N,K", TC_SHAPES)
def test_gemm_tcgen05(lib, name, a_mn, b_mn, M, N, K, backend):
    # operands are exactly TF32-representable, so the only error is fp32 accumulation order
    assert _gemm_case(M, N, K, a_mn, b_mn, backend, 7) < TOL_EXACT
    assert _gemm_case(M, N, K, a_mn, b_mn, backend, 8, bias=True, act=1, resid=True) < TOL_EXACT
    assert _gemm_case(M, N, K, a_mn, b_mn, backend, 9, mask=True) < TOL_EXACT
    assert _gemm_case(M, N, K, a_mn, b_mn, backend, 10, split_k=3) < TOL_EXACT


# the tensor core's fp32 accumulator is not IEEE (aligned / truncated adds): 0.8-1.5e-5 measured on B200
TOL_X3 = 3e-5


@pytest.mark.parametrize("name,a_mn,b_mn", MAJORS)
@pytest.mark.parametrize("M,N,K", TC_SHAPES + [(25600, 1024, 256)])
def test_gemm_tcgen05_3xtf32(lib, name, a_mn, b_mn, M, N, K):
    """backend 2: arbitrary fp32 operands, hi/lo split in shared memory, three MMAs per k-step."""
    if M == 25600 and name != "fwd":
        pytest.skip("full-size schedule checked once")
    assert _gemm_case(M, N, K, a_mn, b_mn, 2, 17) < TOL_X3
    assert _gemm_case(M, N, K, a_mn, b_mn, 2, 18, bias=True, act=1, resid=True) < TOL_X3
    assert _gemm_case(M, N, K, a_mn, b_mn, 2, 19, mask=True) < TOL_X3
    assert _gemm_case(M, N, K, a_mn, b_mn, 2, 20, split_k=3) < TOL_X3
    # tanh of pre-activations with std sqrt(K): absolute pre-activation error ~1e-5*sqrt(K) in the linear region
    assert _gemm_case(M, N, K, a_mn, b_mn, 2, 21, bias=True, act=2) < 2e-4


PAIR_SHAPES = [(256, 256, 32), (512, 256, 64), (200, 1004, 96), (1600, 768, 256), (384, 512, 4000),
               (3000, 1024, 256), (25472, 256, 1024), (300, 128, 160), (1000, 100, 256)]


@pytest.mark.parametrize("name,a_mn,b_mn", MAJORS)
@pytest.mark.parametrize("M,N,K", PAIR_SHAPES)
def test_gemm_tcgen05_pair_3xtf32(lib, name, a_mn, b_mn, M, N, K):
    """backend 5: the 3xTF32 product on CTA pairs (tcgen05 cta_group::2, 256x256 cluster tiles, each CTA
    stages half of the B tile); ragged M / N / K tails, all four operand majors, every epilogue."""
    if M > 20000 and name != "fwd":
        pytest.skip("full-size schedule checked once")
    assert _gemm_case(M, N, K, a_mn, b_mn, 5, 17) < TOL_X3
    assert _gemm_case(M, N, K, a_mn, b_mn, 5, 18, bias=True, act=1, resid=True) < TOL_X3
    assert _gemm_case(M, N, K, a_mn, b_mn, 5, 19, mask=True) < TOL_X3
    assert _gemm_case(M, N, K, a_mn, b_mn, 5, 20, split_k=3) < TOL_X3


@pytest.mark.parametrize("M,N,K", [(256, 256, 3000), (256, 1024, 25472), (512, 768, 1000), (300, 200, 517)])
def test_gemm_pair_fused_bias_gradient(lib, M, N, K):
    """weight-gradient product dW[M,N] += X^T dY with the bias gradient colsum(dY) fused into the same launch
    (summed by the splitter warps from the B tiles of row-block 0, every split-K slice, ragged N and K)."""
    g = torch.Generator().manual_seed(3)
    X = torch.randn(K, M, generator=g)           # stored [K, M]: A = X^T is MN-major
    dY = torch.randn(K, N, generator=g)
    Xd, dYd = dev(X), dev(dY)
    dW = torch.zeros(M, N, device="cuda")
    db = torch.zeros(N, device="cuda")
    gemm(Xd, 1, M, dYd, N, 1, dW, N, M, N, K, 5, accumulate=1, split_k=7, colsum=db)
    torch.cuda.synchronize()
    assert rel_err(dW, X.double().T @ dY.double()) < TOL_X3
    assert rel_err(db, dY.double().sum(0)) < 1e-5


@pytest.mark.parametrize("name,a_mn,b_mn", MAJORS)
@pytest.mark.parametrize("M,N,K", [(256, 256, 32), (512, 256, 64), (200, 1004, 96), (1600, 768, 256), (384, 512, 4000),
                                   (300, 128, 160)])
def test_gemm_pair_mixed_bf16_cross_terms(lib, name, a_mn, b_mn, M, N, K):
    """gemm_mix=1: tf32(A) tf32(B) + bf16(A) bf16(Bl) + bf16(Al) bf16(B); the bf16 operand tiles (64-byte-swizzled,
    K-major and MN-major) are written by the splitter warps.  gemm_mix=2 (bf16x3): bf16(A) bf16(B) + the same two
    cross terms with Xl = X - bf16(X).  Same tolerance as the all-TF32 3-term product."""
    errs = {}
    for mix in (1, 2, 0):       # 0 = all three terms as TF32 MMAs (kept as the reference scheme)
        _lib.check(lib.skf_set_option(b"gemm_mix", mix))
        try:
            errs[mix] = [_gemm_case(M, N, K, a_mn, b_mn, 5, 17),
                         _gemm_case(M, N, K, a_mn, b_mn, 5, 18, bias=True, act=1, resid=True),
                         _gemm_case(M, N, K, a_mn, b_mn, 5, 20, split_k=3)]
        finally:
            _lib.check(lib.skf_set_option(b"gemm_mix", 2))     # back to the default
    assert max(errs[1]) < TOL_X3 and max(errs[2]) < TOL_X3 and max(errs[0]) < TOL_X3, errs


def test_split_bf16_matches_torch(lib):
    """hi = bf16(x) (round to nearest even), lo = bf16(x - hi), bit for bit; odd length exercises the scalar tail."""
    g = torch.Generator().manual_seed(5)
    x = dev(torch.randn(100003, generator=g) * torch.logspace(-6, 6, 100003))
    hi, lo = split_bf16(x)
    torch.cuda.synchronize()
    h = x.bfloat16()
    assert torch.equal(hi, h) and torch.equal(lo, (x - h.float()).bfloat16())
    assert rel_err(hi.float() + lo.float(), x.double()) < 2e-5


@pytest.mark.parametrize("name,a_mn,b_mn", MAJORS)
@pytest.mark.parametrize("M,N,K", [(256, 256, 32), (512, 256, 64), (200, 1000, 96), (1600, 768, 256), (384, 512, 4000),
                                   (300, 128, 160), (260, 1004, 100), (25600, 1024, 256)])
def test_gemm_pair_presplit_weights(lib, name, a_mn, b_mn, M, N, K):
    """bf16x3 with the B operand TMA-loaded from bf16 hi / lo images (Dense forward: B row-major [K, N]; dgrad: B
    K-major), four-stage ring; ragged M / N / K; (260, 1004, 100) has a row pitch that is not a multiple of 8
    elements in one of the majors and must fall back to the in-kernel split."""
    if M > 20000 and name != "fwd":
        pytest.skip("full-size schedule checked once")
    assert _gemm_case(M, N, K, a_mn, b_mn, 5, 17, images=True) < TOL_X3
    assert _gemm_case(M, N, K, a_mn, b_mn, 5, 18, bias=True, act=1, resid=True, images=True) < TOL_X3
    assert _gemm_case(M, N, K, a_mn, b_mn, 5, 19, mask=True, images=True) < TOL_X3


@pytest.mark.parametrize("name,a_mn,b_mn", MAJORS)
@pytest.mark.parametrize("M,N,K", [(25600, 256, 256), (19100, 1004, 64), (19200, 768, 96), (40000, 128, 64)])
def test_gemm_pair_tail_split(lib, name, a_mn, b_mn, M, N, K):
    """More cluster tiles than clusters with a partial last wave (100, 300, 225 and 157 tiles on 74 clusters): the tail
    tiles run as two half-width units (gemm_tail_split, the default).  Every epilogue, weights pre-split or not, a
    ragged last column tile (N = 1004), the 128-column tile variant (N = 128), against float64 — and the same
    launches with the split switched off."""
    if name not in ("fwd", "dgrad") and M != 19200:
        pytest.skip("large schedules checked for the forward / dgrad majors; all majors at one shape")
    for split in (1, 0):
        _lib.check(lib.skf_set_option(b"gemm_tail_split", split))
        try:
            assert _gemm_case(M, N, K, a_mn, b_mn, 5, 17, images=True) < TOL_X3
            assert _gemm_case(M, N, K, a_mn, b_mn, 5, 18, bias=True, act=1, resid=True, images=True) < TOL_X3
            assert _gemm_case(M, N, K, a_mn, b_mn, 5, 19, mask=True) < TOL_X3
        finally:
            _lib.check(lib.skf_set_option(b"gemm_tail_split", 1))


def test_gemm_tcgen05_3xtf32_beats_single_pass(lib):
    g = torch.Generator().manual_seed(13)
    for (M, N, K) in ((512, 256, 256), (300, 1024, 256), (256, 512, 1000)):
        A = torch.randn(M, K, generator=g); Bm = torch.randn(K, N, generator=g)
        Y = torch.zeros(M, N, device="cuda")
        gemm(dev(A), K, 1, dev(Bm), N, 1, Y, N, M, N, K, 2)
        e3 = rel_err(Y, A.double() @ Bm.double())
        gemm(dev(A), K, 1, dev(Bm), N, 1, Y, N, M, N, K, 0)
        e1 = rel_err(Y, A.double() @ Bm.double())
        assert e3 < TOL_X3 and e1 < TOL_TF32 and e3 < e1 / 10, (e1, e3)


def test_gemm_tcgen05_unrounded_operands_within_tolerance(lib):
    """fp32 operands that are NOT pre-rounded get truncated by kind::tf32: still inside 1e-3."""
    g = torch.Generator().manual_seed(12)
    M, N, K = 512, 256, 256
    A = torch.randn(M, K, generator=g); Bm = torch.randn(K, N, generator=g)
    Y = torch.zeros(M, N, device="cuda")
    gemm(dev(A), K, 1, dev(Bm), N, 1, Y, N, M, N, K, 0)
    assert rel_err(Y, A.double() @ Bm.double()) < TOL_TF32


@pytest.mark.parametrize("B,T,E,N", [(3, 7, 16, 128), (5, 33, 64, 256), (4, 200, 32, 512), (2, 19, 16, 1024)])
def test_kv_expand_rank2_form(lib, B, T, E, N):
    """Cross-attention K/V of the expanded bottleneck (sketchformer.py:170-181, transformer.py:370-376, 155-156):
    Dense(N)(emb[b,:] w[t] + beta[t]) == w[t] (emb W)[b,:] + beta[t] colsum(W) + bias, forward and backward."""
    g = torch.Generator().manual_seed(B * T)
    emb = torch.randn(B, E, generator=g).double().requires_grad_(True)
    Wm = (torch.randn(E, N, generator=g) * 0.2).double().requires_grad_(True)
    bias = torch.randn(N, generator=g).double().requires_grad_(True)
    w = torch.randn(T, generator=g).double().requires_grad_(True)
    beta = torch.randn(T, generator=g).double().requires_grad_(True)
    pre = emb[:, None, :] * w[None, :, None] + beta[None, :, None]          # DenseExpander.call
    kv = pre @ Wm + bias
    G = torch.randn(B, T, N, generator=g)
    kv.backward(G.double())
    U = (emb.detach() @ Wm.detach()); s = Wm.detach().sum(0)
    Ud, sd, wd, bd = dev(U.float()), dev(s.float()), dev(w.detach().float()), dev(beta.detach().float())
    out = torch.empty(B, T, N, device="cuda")
    _lib.check(lib.skf_kv_expand_fwd(ptr(Ud), ptr(sd), ptr(dev(bias.detach().float())), ptr(wd), ptr(bd), ptr(out),
                                     B, T, N, stream()))
    assert rel_err(out, kv.detach()) < TOL_EXACT
    gU = torch.zeros(B, N, device="cuda"); gs = torch.zeros(N, device="cuda"); gb = torch.zeros(N, device="cuda")
    gw = torch.zeros(T, device="cuda"); gbeta = torch.zeros(T, device="cuda")
    Gd = dev(G)
    _lib.check(lib.skf_kv_expand_bwd(ptr(Gd), ptr(Ud), ptr(sd), ptr(wd), ptr(bd), ptr(gU), ptr(gs), ptr(gb), ptr(gw),
                                     ptr(gbeta), B, T, N, stream()))
    assert rel_err(gb, bias.grad) < TOL_EXACT
    assert rel_err(gw, w.grad) < TOL_EXACT and rel_err(gbeta, beta.grad) < TOL_EXACT
    # chain rule through U = emb W and s = colsum(W)
    gW = emb.detach().T @ gU.cpu().double() + gs.cpu().double()[None, :]
    assert rel_err(gW, Wm.grad) < TOL_EXACT
    assert rel_err(gU.cpu().double() @ Wm.detach().T, emb.grad) < TOL_EXACT
    Y = torch.randn(E, N, generator=g)
    Yd = dev(Y)
    _lib.check(lib.skf_add_row_broadcast(ptr(Yd), ptr(gs), E, N, stream()))
    assert rel_err(Yd, Y.double() + gs.cpu().double()[None, :]) < TOL_EXACT


@pytest.mark.parametrize("B,H,Tq,T,dh,masked", [(2, 4, 9, 10, 16, False), (3, 8, 33, 40, 32, True), (2, 8, 199, 200, 32, False),
                                                 (2, 2, 70, 257, 64, True), (1, 8, 40, 512, 64, False)])
def test_cross_attention_on_the_rank2_form(lib, B, H, Tq, T, dh, masked):
    """skf_xattn_fold_fwd / _bwd against MultiHeadAttention's second block computed the long way (fp64): expand,
    Dense keys / values, scaled_dot_product_attention with the additive -1e9 key-padding mask, autograd."""
    g = torch.Generator().manual_seed(7 * T + dh)
    d = H * dh
    q = torch.randn(B, Tq, d, generator=g).double().requires_grad_(True)
    U = torch.randn(B, 2 * d, generator=g).double().requires_grad_(True)
    s = torch.randn(2 * d, generator=g).double().requires_grad_(True)
    bias = torch.randn(2 * d, generator=g).double().requires_grad_(True)
    w = torch.randn(T, generator=g).double().requires_grad_(True)
    beta = torch.randn(T, generator=g).double().requires_grad_(True)
    kv = w[None, :, None] * U[:, None, :] + beta[None, :, None] * s[None, None, :] + bias
    k, v = kv[..., :d], kv[..., d:]
    pad = None
    mask = None
    if masked:
        lens = torch.randint(max(1, T // 3), T + 1, (B,), generator=g)
        pad = (torch.arange(T)[None, :] >= lens[:, None])
        mask = pad[:, None, None, :].double()
    o, _ = O.scaled_dot_product_attention(O.split_heads(q, H), O.split_heads(k, H), O.split_heads(v, H), mask)
    o = o.permute(0, 2, 1, 3).reshape(B, Tq, d)
    do = torch.randn(B, Tq, d, generator=g)
    o.backward(do.double())
    f32 = lambda t: dev(t.detach().float())
    qd, Ud, sd, bd, wd, btd, dod = f32(q), f32(U), f32(s), f32(bias), f32(w), f32(beta), dev(do)
    padd = dev(pad.to(torch.uint8)) if masked else None
    od = torch.empty(B, Tq, d, device="cuda")
    _lib.check(lib.skf_xattn_fold_fwd(ptr(qd), d, ptr(Ud), ptr(sd), ptr(bd), ptr(wd), ptr(btd), ptr(padd), ptr(od), d,
                                      B, H, Tq, T, dh, stream()))
    assert rel_err(od, o.detach()) < 2e-5
    dq = torch.empty_like(qd); gU = torch.full((B, 2 * d), 7.0, device="cuda")      # gU is overwritten, the rest add
    gs = torch.zeros(2 * d, device="cuda"); gb = torch.zeros(2 * d, device="cuda")
    gw = torch.zeros(T, device="cuda"); gbeta = torch.zeros(T, device="cuda")
    _lib.check(lib.skf_xattn_fold_bwd(ptr(qd), d, ptr(Ud), ptr(sd), ptr(bd), ptr(wd), ptr(btd), ptr(padd), ptr(dod), d,
                                      ptr(dq), d, ptr(gU), ptr(gs), ptr(gb), ptr(gw), ptr(gbeta), B, H, Tq, T, dh,
                                      stream()))
    tol = 5e-5
    assert rel_err(dq, q.grad) < tol and rel_err(gU, U.grad) < tol and rel_err(gs, s.grad) < tol
    assert rel_err(gw, w.grad) < tol and rel_err(gbeta, beta.grad) < tol
    assert rel_err(gb[d:], bias.grad[d:]) < tol
    assert float(gb[:d].abs().max()) == 0.0 and float(bias.grad[:d].abs().max()) < 1e-9 * float(bias.grad[d:].abs().max() + 1)


# ---- sublayer entry points (skf_ffn_*, skf_logits_ce_*) against torch autograd in float64 ----------------------------
@pytest.mark.parametrize("images", [True, False])
@pytest.mark.parametrize("M,d,f", [(25600, 256, 1024), (1592, 128, 512), (300, 64, 96)])
def test_ffn_sublayer_fwd_bwd(lib, M, d, f, images):
    """skf_ffn_fwd / skf_ffn_bwd = point_wise_feed_forward_network with its residual
    (builders/layers/transformer.py:194-198, 220-222) and its gradients: cfg2's full size (pair GEMM with the tail
    split), a mid size, and a shape that falls to the single-CTA / SIMT kernels."""
    from sketchformer_b200.ops import ffn_bwd, ffn_fwd
    g = torch.Generator().manual_seed(M + d)
    x = torch.randn(M, d, generator=g).double().requires_grad_()
    resid = torch.randn(M, d, generator=g).double()
    W1 = (torch.randn(d, f, generator=g) / math.sqrt(d)).double().requires_grad_()
    b1 = (0.1 * torch.randn(f, generator=g)).double().requires_grad_()
    W2 = (torch.randn(f, d, generator=g) / math.sqrt(f)).double().requires_grad_()
    b2 = (0.1 * torch.randn(d, generator=g)).double().requires_grad_()
    dy = torch.randn(M, d, generator=g).double()
    h_ref = torch.relu(x @ W1 + b1)
    y_ref = resid + h_ref @ W2 + b2
    f32 = lambda t: dev(t.detach().float())
    xd, W1d, b1d, W2d, b2d = f32(x), f32(W1), f32(b1), f32(W2), f32(b2)
    h, y = ffn_fwd(xd, W1d, b1d, W2d, b2d, resid=f32(resid), images=images)
    torch.cuda.synchronize()
    assert rel_err(h, h_ref.detach()) < TOL_X3 and rel_err(y, y_ref.detach()) < TOL_X3
    # gradients at the device's own activation pattern (ReLU is not differentiable at 0: a forward difference of 3e-5
    # flips ~2e-5 of the hidden units and each flip moves dW1 by that unit's whole contribution)
    mask = (h > 0).cpu()
    assert float((mask != (h_ref > 0)).double().mean()) < 1e-3
    with torch.no_grad():
        dW2_ref, db2_ref = h_ref.T @ dy, dy.sum(0)
        dh_ref = (dy @ W2.T) * mask
        dW1_ref, db1_ref, dx_ref = x.T @ dh_ref, dh_ref.sum(0), dh_ref @ W1.T
    extra = dev(torch.randn(M, d, generator=g))
    dx, dW1, db1, dW2, db2 = ffn_bwd(xd, h, f32(dy), W1d, W2d, dx_resid=extra, images=images)
    torch.cuda.synchronize()
    assert rel_err(dW2, dW2_ref) < 2e-4 and rel_err(db2, db2_ref) < 2e-4
    assert rel_err(dW1, dW1_ref) < 2e-4 and rel_err(db1, db1_ref) < 2e-4
    assert rel_err(dx - extra, dx_ref) < 2e-4


@pytest.mark.parametrize("images", [True, False])
@pytest.mark.parametrize("B,Td,d,V", [(128, 199, 256, 1004), (5, 31, 64, 52)])
def test_logits_ce_sublayer_fwd_bwd(lib, B, Td, d, V, images):
    """skf_logits_ce_fwd / _bwd = final_layer + masked token CE + accuracy (models/sketchformer.py:179, 335-339;
    builders/losses.py:26-41): loss sum and dlogits against float64, argmax / correct count exactly."""
    from sketchformer_b200.ops import logits_ce_bwd, logits_ce_fwd
    g = torch.Generator().manual_seed(B * Td)
    x = torch.randn(B * Td, d, generator=g).double().requires_grad_()
    W = (torch.randn(d, V, generator=g) / math.sqrt(d)).double().requires_grad_()
    b = (0.1 * torch.randn(V, generator=g)).double().requires_grad_()
    tgt = torch.randint(1, V, (B, Td), generator=g)
    tgt[:, Td - Td // 4:] = 0                                      # PAD tail: masked in the loss, counted in the mean
    z = x @ W + b
    ce = torch.nn.functional.cross_entropy(z, tgt.reshape(-1), reduction="none") * (tgt.reshape(-1) != 0)
    scale = 1.0 / (B * Td)
    (ce.sum() * scale).backward()
    f32 = lambda t: dev(t.detach().float())
    xd, Wd = f32(x), f32(W)
    dlog, stats, am = logits_ce_fwd(xd, Wd, f32(b), dev(tgt), scale, images=images)
    torch.cuda.synchronize()
    assert abs(float(stats[0]) - float(ce.sum())) < 1e-4 * float(ce.sum())
    zs = z.detach()
    top2 = zs.topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 1e-3 * (zs.max() - zs.min())            # rows whose argmax is not a near tie
    assert torch.equal(am.cpu().long()[clear], zs.argmax(1)[clear])
    assert abs(float(stats[1]) - float((zs.argmax(1) == tgt.reshape(-1)).sum())) <= float((~clear).sum())
    want_dlog = (torch.softmax(zs, 1) - torch.nn.functional.one_hot(tgt.reshape(-1), V)) * (tgt.reshape(-1, 1) != 0) * scale
    assert rel_err(dlog, want_dlog) < 2e-4
    dx, dW, db = logits_ce_bwd(xd, dlog, Wd, B, Td, images=images)
    torch.cuda.synchronize()
    assert rel_err(dx, x.grad) < 3e-4 and rel_err(dW, W.grad) < 3e-4 and rel_err(db, b.grad) < 3e-4

"""End-to-end parity of the CUDA train step (skf_train_step through the C ABI) against the oracle:
activations, losses, every gradient tensor, and the weights after three optimizer steps.

Tolerances (written here as the north-star asks): relative L2 error per tensor
  * 1e-3 with the tcgen05 tensor-core kernels (north-star: "within 1e-3 fp32 relative"),
  * 2e-4 with the exact-fp32 SIMT backend (which isolates tensor-core rounding from logic errors),
token / class argmax bit-exact.

Two properties of the model make a naive per-tensor comparison ill-posed, so the test handles them
explicitly (both hold for ANY fp32 implementation, the reference's included):
  * ReLU is not differentiable at 0.  A forward perturbation eps flips ~eps of the hidden units and
    moves d(loss)/d(ffn.w1) by ~sqrt(eps) in L2 (measured: forward 4e-5 -> gradient 1.7e-3).  The
    gradients are therefore compared with the oracle evaluated at the engine's own ReLU activation
    pattern (oracle.relu_pattern), and the pattern itself is required to agree with the oracle's on
    all but a 1e-3 fraction of units.
  * softmax is shift-invariant, so the key-projection bias (wk.b) has an identically-zero gradient:
    what any implementation computes there is rounding noise, which Adam normalises to O(lr) steps.
    Those tensors are measured against 1e-4 of the global gradient norm and skipped in the
    "weights after three steps" displacement check.
"""
import os

import numpy as np
import pytest
import torch

from oracle import sketchformer_oracle as O
from oracle.make_golden import CFG1, TINY
from sketchformer_b200 import synth
from sketchformer_b200.engine import Engine, make_cfg
from tests.gpu_util import rel_err

pytestmark = pytest.mark.gpu
DROPOUT_SEED = 20260


def build(cfg_kw, B, backend, precision=2, attn_backend=-1, **flags):
    ocfg = O.OracleConfig(**cfg_kw, **flags)
    c = make_cfg(B, ocfg.max_seq_len, ocfg.d_model, ocfg.dff, ocfg.num_heads, ocfg.num_layers, ocfg.vocab_size,
                 ocfg.n_classes, ocfg.lowerdim, continuous=ocfg.use_continuous_data,
                 blind_decoder_mask=ocfg.blind_decoder_mask,
                 do_classification=ocfg.do_classification, do_reconstruction=ocfg.do_reconstruction,
                 class_weight=ocfg.class_weight, recon_weight=ocfg.recon_weight, gemm_backend=backend,
                 attn_backend=backend if attn_backend < 0 else attn_backend, gemm_precision=precision,
                 dropout_rate=ocfg.dropout_rate, dropout_seed=DROPOUT_SEED, attn_version=ocfg.attn_version,
                 class_buffer_layers=ocfg.class_buffer_layers, class_dropout=ocfg.class_dropout)
    return ocfg, Engine(c)


def engine_activations(eng, ocfg, B, materialised_pre=False, materialised_kv=False):
    """Engine workspace tensors under the oracle's tape names."""
    T, d, L = ocfg.max_seq_len, ocfg.d_model, ocfg.num_layers
    out = {"enc.x0": eng.tensor("enc.x0", (B, T, d))}
    for i in range(L):
        qkv = eng.tensor(f"enc.l{i}.qkv", (B, T, 3 * d))
        out[f"enc.l{i}.mha.q"], out[f"enc.l{i}.mha.k"], out[f"enc.l{i}.mha.v"] = qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:]
        out[f"enc.l{i}.mha.ctx"] = eng.tensor(f"enc.l{i}.ctx", (B, T, d))
        out[f"enc.l{i}.out1"] = eng.tensor(f"enc.l{i}.out1", (B, T, d))
        out[f"enc.l{i}.out2"] = eng.tensor(f"enc.l{i}.out2", (B, T, d))
    out["pool.a"] = eng.tensor("pool.a", (B, T))
    if ocfg.do_reconstruction:
        # the expanded bottleneck ("pre_decoder") is not materialised when the plan uses the rank-2 form of the
        # cross-attention keys / values (kv_fold): its Dense(2d) images mha2.k / mha2.v below are compared instead
        if materialised_pre:
            out["pre_decoder"] = eng.tensor("pre_decoder", (B, T, d))
        out["dec.x0"] = eng.tensor("dec.x0", (B, T - 1, d))
        for i in range(L):
            qkv = eng.tensor(f"dec.l{i}.qkv", (B, T - 1, 3 * d))
            out[f"dec.l{i}.mha1.q"], out[f"dec.l{i}.mha1.k"], out[f"dec.l{i}.mha1.v"] = qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:]
            out[f"dec.l{i}.mha1.ctx"] = eng.tensor(f"dec.l{i}.ctx1", (B, T - 1, d))
            out[f"dec.l{i}.mha2.q"] = eng.tensor(f"dec.l{i}.q2", (B, T - 1, d))
            if materialised_kv:      # not under kv_fold=2 (the default): mha2.ctx below covers the whole block
                kv = eng.tensor(f"dec.l{i}.kv2", (B, T, 2 * d))
                out[f"dec.l{i}.mha2.k"], out[f"dec.l{i}.mha2.v"] = kv[..., :d], kv[..., d:]
            out[f"dec.l{i}.mha2.ctx"] = eng.tensor(f"dec.l{i}.ctx2", (B, T - 1, d))
            for j in (1, 2, 3):
                out[f"dec.l{i}.out{j}"] = eng.tensor(f"dec.l{i}.out{j}", (B, T - 1, d))
    return out


def engine_relu_pattern(eng, ocfg, B):
    """name -> bool (B, T, dff): which hidden units of every FFN were active in the engine's forward."""
    T, f, pat = ocfg.max_seq_len, ocfg.dff, {}
    for i in range(ocfg.num_layers):
        pat[f"enc.l{i}.ffn"] = (eng.tensor(f"enc.l{i}.hff", (B, T, f)) > 0).cpu()
        if ocfg.do_reconstruction:
            pat[f"dec.l{i}.ffn"] = (eng.tensor(f"dec.l{i}.hff", (B, T - 1, f)) > 0).cpu()
    return pat


# cfg2 dimensions (head dim 32: the tcgen05 attention kernels) at a batch the fp64 oracle does in ~1 s
CFG2 = dict(num_layers=4, d_model=256, dff=1024, num_heads=8, lowerdim=256, max_seq_len=200, vocab_size=1004,
            n_classes=345)
CASES = [("tiny", TINY, 4), ("cfg1", CFG1, 8), ("cfg2-b6", CFG2, 6)]


@pytest.fixture(params=[2, 1, 0], ids=["xfold", "kvfold", "kvgemm"])
def kv_fold(request):
    """2 (default): the decoder's second attention block on the rank-2 form of the expanded bottleneck, no K / V;
    1: K / V from that form, regular attention kernels; 0: expander + Dense(2d) GEMM + attention kernels."""
    from sketchformer_b200 import _lib
    lib = _lib.load()
    _lib.check(lib.skf_set_option(b"kv_fold", request.param))
    yield request.param
    _lib.check(lib.skf_set_option(b"kv_fold", 2))


@pytest.mark.parametrize("backend", [1, 0], ids=["simt", "tcgen05"])
@pytest.mark.parametrize("name,cfg_kw,B", CASES, ids=[c[0] for c in CASES])
def test_train_step_matches_oracle(name, cfg_kw, B, backend, kv_fold):
    tol = 2e-4 if backend == 1 else 1e-3
    ocfg, eng = build(cfg_kw, B, backend)
    x, y = synth.token_batch(B, ocfg.max_seq_len, ocfg.n_classes, 1234, vocab=ocfg.vocab_size)
    xt, yt = torch.from_numpy(x), torch.from_numpy(y)
    p = O.init_params(ocfg, 0, torch.float64)
    eng.load_params(p)
    opt = O.AdamTF(p, ocfg.d_model)
    xd, yd = xt.cuda(), yt.cuda()
    eng.train_step(xd, yd, apply_adam=True)
    torch.cuda.synchronize()
    # oracle at the engine's ReLU activation pattern (see module docstring)
    pattern = engine_relu_pattern(eng, ocfg, B)
    own, tape = {}, {}
    with O.relu_pattern(pattern, record=own):
        losses, grads = O.train_step(p, opt, xt, yt, ocfg, tape=tape, dtype=torch.float64)
    flips = sum(int((own[k] != pattern[k]).sum()) for k in pattern)
    units = sum(pattern[k].numel() for k in pattern)
    assert flips <= 1e-3 * units, f"ReLU pattern differs on {flips} of {units} hidden units"

    bad = []
    acts = engine_activations(eng, ocfg, B, materialised_pre=kv_fold == 0, materialised_kv=kv_fold < 2)
    for k, ref in tape.items():
        if k not in acts:
            continue
        e = rel_err(acts[k], ref)
        if not e < tol:
            bad.append((k, e))
    assert not bad, f"activations off: {bad[:8]}"

    m = eng.last_step_metrics()
    for k in ("recon_loss", "class_loss", "total_loss"):
        assert abs(m[k] - losses[k]) <= tol * abs(losses[k]), (k, m[k], losses[k])
    assert m["recon_acc"] == pytest.approx(losses["recon_acc"], abs=1e-12)   # integer counts: exact
    assert m["class_acc"] == pytest.approx(losses["class_acc"], abs=1e-12)

    got = eng.export_grads(grads)
    G = float(torch.sqrt(sum((v.double() ** 2).sum() for v in grads.values())))
    for k, ref in grads.items():
        num = float((got[k].double() - ref.double()).norm())
        den = max(float(ref.double().norm()), 1e-4 * G)
        if not num / den < tol:
            bad.append((k, num / den))
    assert not bad, f"gradients off: {sorted(bad, key=lambda t: -t[1])[:8]}"

    # steps 2 and 3: the optimizer really moves the weights (lr(0) = 0 on step 1)
    p_before = {k: v.clone() for k, v in p.items()}
    # jump into the warm-up ramp (lr ~ 1e-3): at iterations 1..2 the displacement (~1e-7) is below fp32 ulp
    opt.iterations = 4000
    eng.set_step(4000)
    for s in (1, 2):
        losses, _ = O.train_step(p, opt, xt, yt, ocfg, dtype=torch.float64)
        eng.train_step(xd, yd, apply_adam=True)
        m = eng.last_step_metrics()
        assert abs(m["total_loss"] - losses["total_loss"]) <= tol * abs(losses["total_loss"]), (s, m, losses)
    assert eng.step_count() == 4002
    newp = eng.export_params(p)
    for k, ref in p.items():
        if float(grads[k].double().norm()) < 1e-6 * G:
            continue      # identically-zero gradient (key bias): Adam random-walks on rounding noise
        # compare the displacement produced by the optimizer (weights barely move in 2 tiny-lr steps)
        num = float(((newp[k].double() - p_before[k]) - (ref - p_before[k])).norm())
        den = max(float((ref - p_before[k]).norm()), 1e-12)
        # Adam normalises by sqrt(v): elements with |g| ~ rounding noise move by +-lr whatever their size,
        # so the displacement of a whole tensor is only reproducible to ~10% (7% measured on ln gains)
        if not num / den < 1.5e-1:
            bad.append((k, num / den))
    assert not bad, f"weights after 3 steps off: {sorted(bad, key=lambda t: -t[1])[:8]}"
    run = eng.metrics()
    assert run["total_loss"] > 0 and 0 <= run["recon_acc"] <= 1


def test_golden_cfg1_losses(golden_dir):
    """The CUDA path reproduces the committed golden losses of cfg1 (oracle fp64)."""
    g = np.load(os.path.join(golden_dir, "cfg1_tokens.npz"))
    ocfg, eng = build(CFG1, 8, 0)
    eng.load_params(O.init_params(ocfg, 0, torch.float32))
    xd, yd = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["y"]).cuda()
    for s in range(3):
        eng.train_step(xd, yd, True)
        m = eng.last_step_metrics()
        for k in ("recon_loss", "class_loss", "total_loss"):
            want = float(g[f"f64/step{s}/{k}"])
            assert abs(m[k] - want) <= 1e-3 * abs(want), (s, k, m[k], want)


def test_encode_matches_oracle_predict_class():
    ocfg, eng = build(CFG1, 8, 0)
    p = O.init_params(ocfg, 0, torch.float64)
    eng.load_params(p)
    x, _ = synth.token_batch(8, ocfg.max_seq_len, ocfg.n_classes, 77, vocab=ocfg.vocab_size)
    want = O.predict_class(p, torch.from_numpy(x), ocfg, torch.float64)
    eng.encode(torch.from_numpy(x).cuda())
    emb = eng.tensor("embedding", (8, ocfg.d_model))
    assert rel_err(emb, want["embedding"]) < 1e-3
    assert torch.equal(eng.tensor("cls.argmax", dtype=torch.int32).cpu()[:8], want["class"])


ABLATIONS = [dict(do_classification=False), dict(do_reconstruction=False), dict(blind_decoder_mask=False),
             dict(class_weight=0.5, recon_weight=2.0)]
# TINY runs the SIMT attention (head dim 16); the small cfg2-shaped case (head dim 32, T=200) runs the tcgen05
# attention and pair-GEMM kernels through the same flags
CFG2_L1 = dict(num_layers=1, d_model=256, dff=1024, num_heads=8, lowerdim=256, max_seq_len=200, vocab_size=1004,
               n_classes=345)


@pytest.mark.parametrize("name,cfg_kw,B,backend", [("tiny", TINY, 4, 1), ("tiny", TINY, 4, 0), ("cfg2-l1", CFG2_L1, 3, 0)],
                         ids=["tiny-simt", "tiny-tcgen05", "cfg2l1-tcgen05"])
@pytest.mark.parametrize("flags", ABLATIONS, ids=["no-class", "no-recon", "not-blind", "weights"])
def test_ablation_flags(flags, name, cfg_kw, B, backend):
    """do_classification / do_reconstruction / blind_decoder_mask / loss-weight variants of build_model()
    (models/sketchformer.py:63-129) on both backends."""
    tol = 2e-4 if backend == 1 else 1e-3
    ocfg, eng = build(cfg_kw, B, backend, **flags)
    x, y = synth.token_batch(B, ocfg.max_seq_len, ocfg.n_classes, 5, vocab=ocfg.vocab_size)
    xt, yt = torch.from_numpy(x), torch.from_numpy(y)
    p = O.init_params(ocfg, 0, torch.float64)
    eng.load_params(p)
    opt = O.AdamTF(p, ocfg.d_model)
    eng.train_step(xt.cuda(), yt.cuda(), True)
    torch.cuda.synchronize()
    with O.relu_pattern(engine_relu_pattern(eng, ocfg, B)):
        losses, grads = O.train_step(p, opt, xt, yt, ocfg, dtype=torch.float64)
    m = eng.last_step_metrics()
    assert abs(m["total_loss"] - losses["total_loss"]) <= tol * abs(losses["total_loss"]), (flags, m, losses)
    got = eng.export_grads(grads)
    G = float(torch.sqrt(sum((v.double() ** 2).sum() for v in grads.values())))
    bad = []
    for k, ref in grads.items():
        if k.endswith(".wk.b"):
            continue      # identically-zero gradient (softmax is shift-invariant): rounding noise on both sides, whose
                          # size relative to G depends on which loss terms are switched on; test_train_step_matches_oracle
                          # bounds it at the default flags
        num = float((got[k].double() - ref.double()).norm())
        if not num / max(float(ref.double().norm()), 1e-4 * G) < tol:
            bad.append((k, num / max(float(ref.double().norm()), 1e-4 * G)))
    assert not bad, f"{flags}: gradients off: {sorted(bad, key=lambda t: -t[1])[:8]}"


CONT_CASES = [("tiny", TINY, 4), ("cfg4-b4", CFG2, 4)]


@pytest.mark.parametrize("backend", [1, 0], ids=["simt", "tcgen05"])
@pytest.mark.parametrize("name,cfg_kw,B", CONT_CASES, ids=[c[0] for c in CONT_CASES])
def test_continuous_train_step_matches_oracle(name, cfg_kw, B, backend):
    """BASELINE.json configs[3]: use_continuous_data=True — Dense(5->d) input embedding, Dense(5) output
    layer, MSE + 3-way CE loss with the reference's normalisation (builders/losses.py:43-66), pad flags from
    the stroke-5 pad bit (builders/utils.py:38-39)."""
    tol = 2e-4 if backend == 1 else 1e-3
    ocfg, eng = build(cfg_kw, B, backend, use_continuous_data=True)
    x, y = synth.continuous_batch(B, ocfg.max_seq_len, ocfg.n_classes, 4321)
    xt, yt = torch.from_numpy(x), torch.from_numpy(y)
    p = O.init_params(ocfg, 0, torch.float64)
    assert "enc.emb.k" in p and tuple(p["out.k"].shape) == (ocfg.d_model, 5)
    eng.load_params(p)
    opt = O.AdamTF(p, ocfg.d_model)
    eng.train_step(xt.cuda(), yt.cuda(), apply_adam=True)
    torch.cuda.synchronize()
    pattern = engine_relu_pattern(eng, ocfg, B)
    own, tape = {}, {}
    with O.relu_pattern(pattern, record=own):
        losses, grads = O.train_step(p, opt, xt, yt, ocfg, tape=tape, dtype=torch.float64)
    flips = sum(int((own[k] != pattern[k]).sum()) for k in pattern)
    assert flips <= 1e-3 * sum(pattern[k].numel() for k in pattern)
    acts = engine_activations(eng, ocfg, B)
    bad = [(k, rel_err(acts[k], ref)) for k, ref in tape.items() if k in acts and not rel_err(acts[k], ref) < tol]
    assert not bad, f"activations off: {bad[:8]}"
    m = eng.last_step_metrics()
    for k in ("recon_loss", "class_loss", "total_loss"):
        assert abs(m[k] - losses[k]) <= tol * abs(losses[k]), (k, m[k], losses[k])
    assert m["class_acc"] == pytest.approx(losses["class_acc"], abs=1e-12)
    got = eng.export_grads(grads)
    G = float(torch.sqrt(sum((v.double() ** 2).sum() for v in grads.values())))
    for k, ref in grads.items():
        num = float((got[k].double() - ref.double()).norm())
        if not num / max(float(ref.double().norm()), 1e-4 * G) < tol:
            bad.append((k, num / max(float(ref.double().norm()), 1e-4 * G)))
    assert not bad, f"gradients off: {sorted(bad, key=lambda t: -t[1])[:8]}"


def test_golden_tiny_continuous_losses(golden_dir):
    """The CUDA path reproduces the committed golden losses of the continuous fixture (oracle fp64)."""
    g = np.load(os.path.join(golden_dir, "tiny_continuous.npz"))
    ocfg, eng = build(TINY, int(g["x"].shape[0]), 0, use_continuous_data=True)
    eng.load_params(O.init_params(ocfg, 0, torch.float32))
    xd, yd = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["y"]).cuda()
    for s in range(3):
        eng.train_step(xd, yd, True)
        m = eng.last_step_metrics()
        for k in ("recon_loss", "class_loss", "total_loss"):
            want = float(g[f"f64/step{s}/{k}"])
            assert abs(m[k] - want) <= 1e-3 * abs(want), (s, k, m[k], want)


@pytest.mark.parametrize("backend", [1, 0], ids=["simt", "tcgen05"])
@pytest.mark.parametrize("name,cfg_kw,B,cont", [("tiny", TINY, 4, False), ("cfg1", CFG1, 8, False),
                                                ("tiny-continuous", TINY, 4, True)], ids=["tiny", "cfg1", "tiny-cont"])
def test_train_step_with_dropout_matches_oracle(name, cfg_kw, B, cont, backend):
    """dropout_rate=0.1 (the reference default, models/sketchformer.py:31): every Dropout site of
    Encoder/Decoder(.Layer).call is active in the train step.  The device masks are counter-based
    (seed, site, optimizer step, element), regenerated in the backward pass; the oracle is given the same
    masks (oracle/dropout_rng.py), so losses and gradients must agree to the usual tolerance — and a
    second step must draw different masks."""
    from oracle.dropout_rng import DeviceMasks
    tol = 2e-4 if backend == 1 else 1e-3
    ocfg, eng = build(cfg_kw, B, backend, dropout_rate=0.1, use_continuous_data=cont)
    if cont:
        x, y = synth.continuous_batch(B, ocfg.max_seq_len, ocfg.n_classes, 77)
    else:
        x, y = synth.token_batch(B, ocfg.max_seq_len, ocfg.n_classes, 77, vocab=ocfg.vocab_size)
    xt, yt = torch.from_numpy(x), torch.from_numpy(y)
    p = O.init_params(ocfg, 0, torch.float64)
    eng.load_params(p)
    opt = O.AdamTF(p, ocfg.d_model)
    first = None
    for step in (0, 4000):
        opt.iterations = step
        eng.set_step(step)
        eng.train_step(xt.cuda(), yt.cuda(), apply_adam=False)
        torch.cuda.synchronize()
        pattern = engine_relu_pattern(eng, ocfg, B)
        with O.dropout_masks(DeviceMasks(0.1, DROPOUT_SEED, step)), O.relu_pattern(pattern):
            p0 = {k: v.clone() for k, v in p.items()}
            losses, grads = O.train_step(p0, O.AdamTF(p0, ocfg.d_model), xt, yt, ocfg, dtype=torch.float64)
        m = eng.last_step_metrics()
        for k in ("recon_loss", "class_loss", "total_loss"):
            assert abs(m[k] - losses[k]) <= tol * abs(losses[k]), (step, k, m[k], losses[k])
        got = eng.export_grads(grads)
        G = float(torch.sqrt(sum((v.double() ** 2).sum() for v in grads.values())))
        bad = []
        for k, ref in grads.items():
            num = float((got[k].double() - ref.double()).norm())
            if not num / max(float(ref.double().norm()), 1e-4 * G) < tol:
                bad.append((k, num / max(float(ref.double().norm()), 1e-4 * G)))
        assert not bad, f"step {step}: gradients off: {sorted(bad, key=lambda t: -t[1])[:8]}"
        if first is None:
            first = losses["total_loss"]
        else:
            assert abs(first - losses["total_loss"]) > 1e-6 * abs(first), "the mask did not change with the step"
    # inference entry points never drop (training=False)
    ocfg0, eng0 = build(cfg_kw, B, backend, use_continuous_data=cont)
    eng0.load_params(p)
    eng.encode(xt.cuda()); eng0.encode(xt.cuda())
    assert torch.equal(eng.tensor("embedding", (B, ocfg.d_model)), eng0.tensor("embedding", (B, ocfg.d_model)))


@pytest.mark.parametrize("backend", [1, 0], ids=["simt", "tcgen05"])
@pytest.mark.parametrize("flags", [dict(attn_version=2), dict(class_buffer_layers=2),
                                   dict(attn_version=2, class_buffer_layers=1, dropout_rate=0.1, class_dropout=0.2)],
                         ids=["selfattn-v2", "class-buffer-2", "v2+buffer+dropout"])
def test_bottleneck_variants_match_oracle(flags, backend):
    """attn_version=2 (SelfAttnV2: (d,d) pooling attention + Dense(lowerdim), embedding / expander / cross
    keys-values lowerdim wide, transformer.py:80-137) and class_buffer_layers (Dense(lowerdim, relu) +
    Dropout(class_dropout) before the classifier, models/sketchformer.py:99-104, 183-199)."""
    from oracle.dropout_rng import DeviceMasks
    tol = 2e-4 if backend == 1 else 1e-3
    B = 4
    ocfg, eng = build(TINY, B, backend, **flags)
    x, y = synth.token_batch(B, ocfg.max_seq_len, ocfg.n_classes, 31, vocab=ocfg.vocab_size)
    xt, yt = torch.from_numpy(x), torch.from_numpy(y)
    p = O.init_params(ocfg, 0, torch.float64)
    eng.load_params(p)
    eng.train_step(xt.cuda(), yt.cuda(), apply_adam=False)
    torch.cuda.synchronize()
    pattern = engine_relu_pattern(eng, ocfg, B)
    masks = DeviceMasks(ocfg.dropout_rate, DROPOUT_SEED, 0, class_rate=ocfg.class_dropout)
    with O.dropout_masks(masks), O.relu_pattern(pattern):
        losses, grads = O.train_step(p, O.AdamTF(p, ocfg.d_model), xt, yt, ocfg, dtype=torch.float64)
    m = eng.last_step_metrics()
    for k in ("recon_loss", "class_loss", "total_loss"):
        assert abs(m[k] - losses[k]) <= tol * abs(losses[k]), (k, m[k], losses[k])
    got = eng.export_grads(grads)
    G = float(torch.sqrt(sum((v.double() ** 2).sum() for v in grads.values())))
    bad = []
    for k, ref in grads.items():
        num = float((got[k].double() - ref.double()).norm())
        if not num / max(float(ref.double().norm()), 1e-4 * G) < tol:
            bad.append((k, num / max(float(ref.double().norm()), 1e-4 * G)))
    assert not bad, f"gradients off: {sorted(bad, key=lambda t: -t[1])[:8]}"
    # inference: embedding width, class prediction and greedy decode through the same variants
    E = ocfg.d_model if ocfg.attn_version == 1 else ocfg.lowerdim
    want = O.predict(p, xt, ocfg, sos=ocfg.vocab_size - 2, eos=ocfg.vocab_size - 1, dtype=torch.float64)
    eng.encode(xt.cuda())
    assert rel_err(eng.tensor("embedding", (B, E)), want["embedding"]) < tol
    assert torch.equal(eng.tensor("cls.argmax", dtype=torch.int32).cpu()[:B], want["class"])
    got_tok = eng.decode(None, None, ocfg.vocab_size - 2, ocfg.vocab_size - 1).cpu().numpy()
    from tests.test_gpu_decode import _assert_same_tokens
    _assert_same_tokens(p, ocfg, want["embedding"], got_tok, want["recon"].numpy())


@pytest.mark.parametrize("backend", [1, 0], ids=["simt", "tcgen05"])
def test_no_bottleneck_matches_oracle(backend):
    """lowerdim=0 (models/sketchformer.py:96, 157-158, 175-176): no pooling / classifier / expander; the decoder's
    cross-attention reads the encoder output and its gradient flows back into the encoder through it."""
    tol = 2e-4 if backend == 1 else 1e-3
    B = 4
    ocfg, eng = build(dict(TINY, lowerdim=0), B, backend, do_classification=False)
    x, y = synth.token_batch(B, ocfg.max_seq_len, ocfg.n_classes, 41, vocab=ocfg.vocab_size)
    xt, yt = torch.from_numpy(x), torch.from_numpy(y)
    p = O.init_params(ocfg, 0, torch.float64)
    assert "pool.W" not in p and "cls.k" not in p and "exp.k" not in p
    eng.load_params(p)
    eng.train_step(xt.cuda(), yt.cuda(), apply_adam=False)
    torch.cuda.synchronize()
    with O.relu_pattern(engine_relu_pattern(eng, ocfg, B)):
        losses, grads = O.train_step(p, O.AdamTF(p, ocfg.d_model), xt, yt, ocfg, dtype=torch.float64)
    m = eng.last_step_metrics()
    assert abs(m["total_loss"] - losses["total_loss"]) <= tol * abs(losses["total_loss"])
    got = eng.export_grads(grads)
    G = float(torch.sqrt(sum((v.double() ** 2).sum() for v in grads.values())))
    bad = []
    for k, ref in grads.items():
        num = float((got[k].double() - ref.double()).norm())
        if not num / max(float(ref.double().norm()), 1e-4 * G) < tol:
            bad.append((k, num / max(float(ref.double().norm()), 1e-4 * G)))
    assert not bad, f"gradients off: {sorted(bad, key=lambda t: -t[1])[:8]}"
    want = O.predict(p, xt, ocfg, sos=ocfg.vocab_size - 2, eos=ocfg.vocab_size - 1, dtype=torch.float64)
    eng.encode(xt.cuda())
    got_tok = eng.decode(None, None, ocfg.vocab_size - 2, ocfg.vocab_size - 1).cpu().numpy()
    from tests.test_gpu_decode import _assert_same_tokens
    _assert_same_tokens(p, ocfg, want["embedding"], got_tok, want["recon"].numpy())


def test_cfg5_dimensions_match_oracle():
    """BASELINE.json configs[4] dimensions (d_model=512, dff=2048, 8 heads of 64, max_seq_len=512) at 2 layers and
    batch 2 so that the fp64 oracle finishes in seconds: head dim 64 and T = 512 run the tcgen05 attention kernels
    (one CTA per SM at this head dim), the GEMMs the tcgen05 pair kernel with 512 / 2048-wide tiles."""
    kw = dict(num_layers=2, d_model=512, dff=2048, num_heads=8, lowerdim=256, max_seq_len=512, vocab_size=1004,
              n_classes=345)
    B = 2
    ocfg, eng = build(kw, B, 0)
    x, y = synth.token_batch(B, ocfg.max_seq_len, ocfg.n_classes, 55, vocab=ocfg.vocab_size)
    xt, yt = torch.from_numpy(x), torch.from_numpy(y)
    p = O.init_params(ocfg, 0, torch.float64)
    eng.load_params(p)
    eng.train_step(xt.cuda(), yt.cuda(), apply_adam=False)
    torch.cuda.synchronize()
    with O.relu_pattern(engine_relu_pattern(eng, ocfg, B)):
        losses, grads = O.train_step(p, O.AdamTF(p, ocfg.d_model), xt, yt, ocfg, dtype=torch.float64)
    m = eng.last_step_metrics()
    for k in ("recon_loss", "class_loss", "total_loss"):
        assert abs(m[k] - losses[k]) <= 1e-3 * abs(losses[k]), (k, m[k], losses[k])
    got = eng.export_grads(grads)
    G = float(torch.sqrt(sum((v.double() ** 2).sum() for v in grads.values())))
    bad = []
    for k, ref in grads.items():
        num = float((got[k].double() - ref.double()).norm())
        if not num / max(float(ref.double().norm()), 1e-4 * G) < 1e-3:
            bad.append((k, num / max(float(ref.double().norm()), 1e-4 * G)))
    assert not bad, f"gradients off: {sorted(bad, key=lambda t: -t[1])[:8]}"

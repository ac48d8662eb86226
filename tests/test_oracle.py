"""CPU tests of the oracle (the restatement of the reference's TF-2.1 graph).

The reference ships no tests or golden vectors, so the oracle is pinned by
(1) its own frozen outputs in tests/golden (drift guard), (2) fp64-vs-fp32 agreement,
(3) finite differences, (4) independent torch.nn.functional implementations of each block and
(5) the closed-form facts of SURVEY §8c (lr(0)=0, masked keys get exactly zero weight, ...).
"""
import math
import os
import zlib

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import sketchformer_oracle as O
from oracle.make_golden import CFG1, TINY, TINY_LOWERDIM0, TINY_NOTBLIND, TINY_V2
from sketchformer_b200 import synth


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.mark.parametrize("name,kw,cont,B", [("tiny_tokens.npz", TINY, False, 4),
                                            ("tiny_continuous.npz", TINY, True, 4),
                                            ("cfg1_tokens.npz", CFG1, False, 8),
                                            ("tiny_v2.npz", TINY_V2, False, 4),
                                            ("tiny_lowerdim0.npz", TINY_LOWERDIM0, False, 4),
                                            ("tiny_notblind.npz", TINY_NOTBLIND, False, 4)])
def test_golden_replay(golden_dir, name, kw, cont, B):
    g = _load(golden_dir, name)
    cfg = O.OracleConfig(use_continuous_data=cont, **kw)
    gen = synth.continuous_batch if cont else synth.token_batch
    x, y = (gen(B, cfg.max_seq_len, cfg.n_classes, 1234) if cont
            else gen(B, cfg.max_seq_len, cfg.n_classes, 1234, vocab=cfg.vocab_size))
    assert np.array_equal(x, g["x"]) and np.array_equal(y, g["y"])
    p = O.init_params(cfg, 0, torch.float32)
    crc = np.array([zlib.crc32(np.ascontiguousarray(v.numpy()).tobytes()) for v in p.values()], dtype=np.uint32)
    assert np.array_equal(crc, g["param_crc"]), "torch CPU RNG stream changed: regenerate goldens"
    p = O.init_params(cfg, 0, torch.float64)
    opt = O.AdamTF(p, cfg.d_model)
    for s in range(3):
        tape = {} if s == 0 else None
        losses, grads = O.train_step(p, opt, torch.from_numpy(x), torch.from_numpy(y), cfg, tape=tape,
                                     dtype=torch.float64)
        for k, v in losses.items():
            assert v == pytest.approx(float(g[f"f64/step{s}/{k}"]), rel=1e-9, abs=1e-12)
        if s == 0:
            for k, v in grads.items():
                assert float(v.norm()) == pytest.approx(float(g[f"grad_norm/{k}"]), rel=1e-7, abs=1e-12)
            for k, v in tape.items():
                assert float(v.norm()) == pytest.approx(float(g[f"act_norm/{k}"]), rel=1e-7)


def test_fp32_tracks_fp64(golden_dir):
    g = _load(golden_dir, "cfg1_tokens.npz")
    for s in range(3):
        for k in ("recon_loss", "class_loss", "total_loss"):
            a, b = float(g[f"f32/step{s}/{k}"]), float(g[f"f64/step{s}/{k}"])
            assert abs(a - b) <= 2e-5 * abs(b)


def test_first_step_has_zero_lr(golden_dir):
    # lr(0)=0 => step 0 and step 1 see identical weights => identical losses
    g = _load(golden_dir, "tiny_tokens.npz")
    assert float(g["f64/step0/total_loss"]) == float(g["f64/step1/total_loss"])
    assert float(g["f64/step2/total_loss"]) != float(g["f64/step1/total_loss"])
    assert O.warmup_decay_lr(0, 256) == 0.0
    assert O.warmup_decay_lr(5000, 256) == pytest.approx(256 ** -0.5 * 5000 ** -0.5)
    assert O.warmup_decay_lr(1, 128) == pytest.approx(128 ** -0.5 * 5000 ** -1.5)
    assert O.warmup_decay_lr(20000, 128) == pytest.approx(128 ** -0.5 * 20000 ** -0.5)


def test_param_counts():
    # SURVEY §8d
    assert O.count_params(O.init_params(O.OracleConfig(**CFG1))) == 2272630
    cfg2 = O.OracleConfig(d_model=256, dff=1024, max_seq_len=200, n_classes=345)
    assert O.count_params(O.init_params(cfg2)) == 8299989
    cfg4 = O.OracleConfig(d_model=256, dff=1024, max_seq_len=200, n_classes=345, use_continuous_data=True)
    assert O.count_params(O.init_params(cfg4)) == 7532270


def test_positional_encoding_and_masks():
    pe = O.positional_encoding(1000, 128)
    assert pe.shape == (1, 1000, 128) and pe.dtype == torch.float32
    assert pe[0, 0, 0] == 0 and pe[0, 0, 1] == 1
    assert float(pe[0, 3, 0]) == pytest.approx(math.sin(3.0), rel=1e-6)
    assert float(pe[0, 7, 5]) == pytest.approx(math.cos(7 / 10000 ** (4 / 128)), rel=1e-6)
    x = torch.tensor([[5, 6, 0, 0], [7, 0, 0, 0]])
    enc, comb, dec = O.create_masks(x, x[:, :-1], torch.float32)
    assert enc.shape == (2, 1, 1, 4) and comb.shape == (2, 1, 3, 3)
    assert enc[0, 0, 0].tolist() == [0, 0, 1, 1]
    assert comb[0, 0].tolist() == [[0, 1, 1], [0, 0, 1], [0, 0, 1]]
    assert comb[1, 0].tolist() == [[0, 1, 1], [0, 1, 1], [0, 1, 1]]
    xc = torch.zeros(1, 3, 5); xc[0, 2, 4] = 1
    assert O.create_padding_mask(xc, torch.float32)[0, 0, 0].tolist() == [0, 0, 1]


def test_sdpa_against_torch_functional():
    g = torch.Generator().manual_seed(1)
    q = torch.randn(2, 4, 7, 16, generator=g, dtype=torch.float64)
    k = torch.randn(2, 4, 9, 16, generator=g, dtype=torch.float64)
    v = torch.randn(2, 4, 9, 16, generator=g, dtype=torch.float64)
    pad = torch.zeros(2, 1, 1, 9, dtype=torch.float64); pad[0, ..., 6:] = 1; pad[1, ..., 3:] = 1
    o, w = O.scaled_dot_product_attention(q, k, v, pad)
    ref = F.scaled_dot_product_attention(q, k, v, attn_mask=(pad == 0))
    assert torch.allclose(o, ref, atol=1e-12)
    assert torch.allclose(w.sum(-1), torch.ones_like(w.sum(-1)), atol=1e-12)
    assert float(w[0, :, :, 6:].abs().max()) == 0.0 and float(w[1, :, :, 3:].abs().max()) == 0.0
    # fully-masked row in fp32 -> uniform, not NaN
    full = torch.ones(1, 1, 1, 9)
    _, wf = O.scaled_dot_product_attention(q[:1].float(), k[:1].float(), v[:1].float(), full)
    assert torch.allclose(wf, torch.full_like(wf, 1 / 9), atol=1e-6)


def _load_mha(m, p, name):
    """oracle Dense kernels are (in, out); torch.nn.MultiheadAttention packs (out, in) rows of q | k | v."""
    with torch.no_grad():
        m.in_proj_weight.copy_(torch.cat([p[name + ".wq.k"].T, p[name + ".wk.k"].T, p[name + ".wv.k"].T]))
        m.in_proj_bias.copy_(torch.cat([p[name + ".wq.b"], p[name + ".wk.b"], p[name + ".wv.b"]]))
        m.out_proj.weight.copy_(p[name + ".wo.k"].T)
        m.out_proj.bias.copy_(p[name + ".wo.b"])


def _load_ffn_ln(layer, p, n, nln):
    with torch.no_grad():
        layer.linear1.weight.copy_(p[n + ".ffn.w1.k"].T); layer.linear1.bias.copy_(p[n + ".ffn.w1.b"])
        layer.linear2.weight.copy_(p[n + ".ffn.w2.k"].T); layer.linear2.bias.copy_(p[n + ".ffn.w2.b"])
        for i in range(1, nln + 1):
            getattr(layer, f"norm{i}").weight.copy_(p[f"{n}.ln{i}.g"])
            getattr(layer, f"norm{i}").bias.copy_(p[f"{n}.ln{i}.b"])


def _randomise(p, seed):
    """Keras initialisers leave biases 0 and LayerNorm at (1, 0): perturb everything so that no term hides."""
    g = torch.Generator().manual_seed(seed)
    for k in p:
        if k.endswith(".b") or k.endswith(".g"):
            p[k] = p[k] + 0.1 * torch.randn(p[k].shape, generator=g, dtype=p[k].dtype)
    return p


def test_encoder_and_decoder_stacks_against_torch_nn_transformer_layers():
    """Independent pin of the layer wiring (the reference cannot run here): the oracle's Encoder / Decoder stacks
    (builders/layers/transformer.py:215-224, 245-262: post-LN residual blocks, LayerNorm eps 1e-6, ReLU FFN,
    additive -1e9 masks) against torch.nn.TransformerEncoderLayer / TransformerDecoderLayer carrying the same
    weights, in float64, on a padded batch."""
    cfg = O.OracleConfig(**TINY)
    d, H, f, L = cfg.d_model, cfg.num_heads, cfg.dff, cfg.num_layers
    p = _randomise(O.init_params(cfg, 3, torch.float64), 4)
    x, _ = synth.token_batch(5, cfg.max_seq_len, cfg.n_classes, 77, vocab=cfg.vocab_size)
    x = torch.from_numpy(x)
    tar = x[:, :-1]
    enc_mask, combined, _ = O.create_masks(x, tar, torch.float64)
    assert bool((x == 0).any()), "the batch must contain padding"

    # ---- encoder
    got = O.encoder(p, x, enc_mask, cfg, False, None, None, torch.float64)
    h = O.embed_input(p, "enc", x, cfg, torch.float64)
    for i in range(L):
        lay = torch.nn.TransformerEncoderLayer(d, H, f, dropout=0.0, activation="relu", layer_norm_eps=1e-6,
                                               batch_first=True, norm_first=False, dtype=torch.float64)
        lay.train()                      # no inference fast path (it zeroes padded rows); dropout is 0
        _load_mha(lay.self_attn, p, f"enc.l{i}.mha")
        _load_ffn_ln(lay, p, f"enc.l{i}", 2)
        h = lay(h, src_key_padding_mask=(x == 0))
    assert torch.allclose(got, h, atol=1e-10), float((got - h).abs().max())

    # ---- decoder: causal + target padding on block 1, memory padding on block 2
    g = torch.Generator().manual_seed(9)
    memory = torch.randn(5, cfg.max_seq_len, d, generator=g, dtype=torch.float64)
    for mem_mask in (torch.zeros_like(enc_mask), enc_mask):        # blind_decoder_mask True / False
        got, _ = O.decoder(p, tar, memory, combined, mem_mask, cfg, False, None, None, torch.float64)
        h = O.embed_input(p, "dec", tar, cfg, torch.float64)
        Td = tar.shape[1]
        causal = torch.triu(torch.ones(Td, Td, dtype=torch.bool), diagonal=1)
        for i in range(L):
            lay = torch.nn.TransformerDecoderLayer(d, H, f, dropout=0.0, activation="relu", layer_norm_eps=1e-6,
                                                   batch_first=True, norm_first=False, dtype=torch.float64)
            lay.train()
            _load_mha(lay.self_attn, p, f"dec.l{i}.mha1")
            _load_mha(lay.multihead_attn, p, f"dec.l{i}.mha2")
            _load_ffn_ln(lay, p, f"dec.l{i}", 3)
            h = lay(h, memory, tgt_mask=causal, tgt_key_padding_mask=(tar == 0),
                    memory_key_padding_mask=(mem_mask[:, 0, 0, :] > 0))
        # rows whose every key is masked (a PAD query never is: key 0 is SOS) do not occur; compare everything
        assert torch.allclose(got, h, atol=1e-10), float((got - h).abs().max())


def test_layernorm_and_losses_against_functional():
    g = torch.Generator().manual_seed(2)
    x = torch.randn(3, 5, 32, generator=g, dtype=torch.float64)
    p = {"ln.g": torch.randn(32, generator=g, dtype=torch.float64), "ln.b": torch.randn(32, generator=g, dtype=torch.float64)}
    assert torch.allclose(O.layer_norm(p, "ln", x), F.layer_norm(x, (32,), p["ln.g"], p["ln.b"], eps=1e-6), atol=1e-12)
    logits = torch.randn(3, 5, 11, generator=g, dtype=torch.float64)
    tgt = torch.tensor([[1, 2, 3, 0, 0], [4, 0, 0, 0, 0], [5, 6, 7, 8, 9]])
    ref = F.cross_entropy(logits.reshape(-1, 11), tgt.reshape(-1), ignore_index=0, reduction="sum") / 15
    assert float(O.reconstruction_loss(tgt, logits)) == pytest.approx(float(ref), rel=1e-12)
    cl = torch.randn(4, 10, generator=g, dtype=torch.float64)
    lab = torch.tensor([[1], [0], [9], [3]])
    assert float(O.class_loss(lab, cl)) == pytest.approx(float(F.cross_entropy(cl, lab[:, 0])), rel=1e-12)


def test_continuous_loss_formula():
    # SURVEY A.4, written out with explicit loops
    g = torch.Generator().manual_seed(3)
    x, _ = synth.continuous_batch(2, 6, 3, 5)
    real = torch.from_numpy(x).double()
    pred = torch.randn(2, 6, 5, generator=g, dtype=torch.float64)
    B, T = 2, 6
    M = 0.0
    for b in range(B):
        for t in range(T):
            z = pred[b, t, 2:]
            M += float(torch.logsumexp(z, 0) - z[int(real[b, t, 2:].argmax())])
    M /= B * T
    tot = 0.0
    for b in range(B):
        for t in range(T):
            loc = 0.5 * float(((real[b, t, :2] - pred[b, t, :2]) ** 2).sum())
            tot += (loc + M) * (0.0 if real[b, t, 4] == 1 else 1.0)
    assert float(O.continuous_reconstruction_loss(real, pred)) == pytest.approx(tot / (B * T), rel=1e-12)


def test_dense_expander_is_rank1():
    p = {"exp.k": torch.tensor([[1., 2., 3.]]), "exp.b": torch.tensor([.5, .25, 0.])}
    e = torch.tensor([[1., -1.], [2., 0.]])
    pre = O.dense_expander(p, e)
    assert pre.shape == (2, 3, 2)
    for b in range(2):
        for t in range(3):
            for c in range(2):
                assert float(pre[b, t, c]) == float(e[b, c] * p["exp.k"][0, t] + p["exp.b"][t])


def test_finite_difference_gradients():
    cfg = O.OracleConfig(num_layers=1, d_model=16, dff=32, num_heads=2, lowerdim=8, max_seq_len=6,
                         vocab_size=12, n_classes=3)
    x, y = synth.token_batch(3, 6, 3, 7, vocab=12)
    xt, yt = torch.from_numpy(x), torch.from_numpy(y)
    p = O.init_params(cfg, 3, torch.float64)
    for v in p.values():
        v.requires_grad_(True)
    _, losses = O.forward_losses(p, xt, yt, cfg, True, None, None, torch.float64)
    losses["total_loss"].backward()
    rng = np.random.default_rng(0)
    for name in ("enc.emb", "enc.l0.mha.wk.k", "enc.l0.ln1.g", "pool.V", "exp.k", "dec.l0.mha2.wv.k",
                 "dec.l0.ffn.w1.b", "out.k", "cls.b"):
        w = p[name]
        flat = w.detach().reshape(-1)
        for idx in rng.choice(flat.numel(), size=min(3, flat.numel()), replace=False):
            old = float(flat[idx]); h = 1e-6
            with torch.no_grad():
                w.reshape(-1)[idx] = old + h
                lp = float(O.forward_losses(p, xt, yt, cfg, True, None, None, torch.float64)[1]["total_loss"])
                w.reshape(-1)[idx] = old - h
                lm = float(O.forward_losses(p, xt, yt, cfg, True, None, None, torch.float64)[1]["total_loss"])
                w.reshape(-1)[idx] = old
            fd = (lp - lm) / (2 * h)
            assert float(w.grad.reshape(-1)[idx]) == pytest.approx(fd, rel=2e-5, abs=2e-8), name


def test_adam_matches_tf_form():
    w = {"a": torch.tensor([1.0, -2.0], dtype=torch.float64)}
    opt = O.AdamTF(w, 128)
    opt.iterations = 10
    g = {"a": torch.tensor([0.5, 0.25], dtype=torch.float64)}
    lr = O.warmup_decay_lr(10, 128)
    m = 0.1 * g["a"]; v = 0.02 * g["a"] ** 2
    alpha = lr * math.sqrt(1 - 0.98 ** 11) / (1 - 0.9 ** 11)
    expect = torch.tensor([1.0, -2.0], dtype=torch.float64) - alpha * m / (v.sqrt() + 1e-9)
    opt.apply(w, g)
    assert torch.allclose(w["a"], expect, atol=1e-15) and opt.iterations == 11


def test_greedy_decode_runs_and_starts_with_sos(golden_dir):
    g = _load(golden_dir, "tiny_tokens.npz")
    rec = g["pred/recon"]
    assert rec.shape[0] == 2 and rec.shape[1] <= TINY["max_seq_len"] + 1
    assert (rec[:, 0] == TINY["vocab_size"] - 2).all()


def test_oracle_matches_vectors_made_by_the_reference_model_code(golden_dir):
    """tests/golden/reference_model_tiny.npz holds what the reference's own models/sketchformer.py + builders/* code
    computed (oracle/make_reference_golden.py runs it from /root/reference on oracle/tf_shim.py): forward outputs,
    both losses and the greedy decode for five graph-changing hparam variants.  This test needs no /root/reference."""
    from oracle.make_reference_golden import VARIANTS, inputs, params
    g = np.load(os.path.join(golden_dir, "reference_model_tiny.npz"))
    for name, kw in VARIANTS:
        cfg = O.OracleConfig(**kw)
        p = params(cfg)
        x, y = inputs(cfg)
        assert np.array_equal(x.numpy(), g[f"{name}/x"]) and np.array_equal(y.numpy(), g[f"{name}/y"])
        tar_inp, tar_real = x[:, :-1], x[:, 1:]
        out = O.call(p, x, tar_inp, cfg, False, O.create_masks(x, tar_inp, torch.float64), dtype=torch.float64)
        assert np.allclose(out["embedding"].numpy(), g[f"{name}/embedding"], atol=1e-12), name
        assert np.allclose(out["recon"].numpy(), g[f"{name}/recon"], atol=1e-11), name
        lo = O.compute_losses(out, tar_real, y, cfg)
        assert float(lo["recon_loss"]) == pytest.approx(float(g[f"{name}/recon_loss"]), rel=1e-12)
        if f"{name}/class" in g:
            assert np.allclose(out["class"].numpy(), g[f"{name}/class"], atol=1e-12)
            assert float(lo["class_loss"]) == pytest.approx(float(g[f"{name}/class_loss"]), rel=1e-9)
        if f"{name}/greedy" in g:
            pr = O.predict(p, x, cfg, sos=cfg.vocab_size - 2, eos=cfg.vocab_size - 1, dtype=torch.float64)
            assert pr["recon"].numpy().tolist() == g[f"{name}/greedy"].tolist(), name
            if f"{name}/pred_class" in g:
                assert pr["class"].numpy().tolist() == g[f"{name}/pred_class"].tolist()


def test_train_steps_match_vectors_made_by_the_reference_trainer(golden_dir):
    """Three steps of the reference's model_trainer / train_on_batch (frozen by oracle/make_reference_golden.py from
    the reference's code on oracle/tf_shim.py): running-mean quick metrics after every step and the logits of the
    updated model.  Needs no /root/reference."""
    from oracle.make_reference_golden import TRAIN_KW, inputs, params, train_inputs
    g = np.load(os.path.join(golden_dir, "reference_model_tiny.npz"))
    cfg = O.OracleConfig(**TRAIN_KW)
    p = params(cfg)
    opt = O.AdamTF(p, cfg.d_model)
    run = O.RunningMetrics()
    for step in range(3):
        xb, yb = train_inputs(cfg, step)
        losses, _ = O.train_step(p, opt, xb, yb, cfg, dtype=torch.float64)
        run.update(losses)
        for k, v in run.result().items():
            assert float(v) == pytest.approx(float(g[f"train/step{step}/{k}"]), rel=1e-6, abs=1e-9), (step, k)
    x, _ = inputs(cfg)
    out = O.call(p, x, x[:, :-1], cfg, False, O.create_masks(x, x[:, :-1], torch.float64), dtype=torch.float64)
    assert np.allclose(out["recon"].detach().numpy(), g["train/recon_after"], atol=1e-9)

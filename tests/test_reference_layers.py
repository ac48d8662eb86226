"""The oracle against the REFERENCE'S OWN layer code.

oracle/tf_shim.py executes builders/utils.py, builders/schedulers.py and builders/layers/transformer.py from
/root/reference on a torch-backed stand-in for the few dozen TensorFlow primitives they call (TensorFlow itself cannot
be installed here).  These tests load the oracle's parameters into the reference's Encoder / Decoder / SelfAttnV1 /
SelfAttnV2 / DenseExpander / MultiHeadAttention objects and require identical outputs in float64: the oracle's wiring
(argument order, masks, residual / LayerNorm order, scalings, transposes, softmax axes) is the reference's.
They are skipped where /root/reference is absent (the GPU box)."""
import math

import numpy as np
import pytest
import torch

from oracle import sketchformer_oracle as O
from oracle import tf_shim
from oracle.reference_model import randomise as _randomise, set_dense as _set_dense, set_ffn as _set_ffn, set_ln as _set_ln, set_mha as _set_mha
from oracle.make_golden import TINY
from sketchformer_b200 import synth

pytestmark = pytest.mark.skipif(not tf_shim.available(), reason="/root/reference is not present on this box")
F64 = torch.float64


@pytest.fixture(scope="module")
def ref():
    return tf_shim.load_reference_builders()


def _batch(cfg, B=5, seed=77):
    x, _ = synth.token_batch(B, cfg.max_seq_len, cfg.n_classes, seed, vocab=cfg.vocab_size)
    x = torch.from_numpy(x)
    assert bool((x == 0).any())
    return x


def test_positional_encoding_masks_and_attention(ref):
    utils, _, _ = ref
    pe = utils.positional_encoding(1000, 64)
    assert pe.dtype == torch.float32 and torch.equal(pe, O.positional_encoding(1000, 64))
    cfg = O.OracleConfig(**TINY)
    x = _batch(cfg)
    tar = x[:, :-1]
    want = O.create_masks(x, tar, torch.float32)
    got = utils.create_masks(x, tar)
    for a, b in zip(got, want):
        assert a.shape == b.shape and torch.equal(a.to(torch.float32), b)
    xc = torch.zeros(2, 6, 5); xc[0, 4:, 4] = 1; xc[1, 2:, 4] = 1          # stroke-5 rows: pad flag is the last bit
    assert torch.equal(utils.create_padding_mask(xc).to(torch.float32), O.create_padding_mask(xc, torch.float32))
    g = torch.Generator().manual_seed(1)
    q, k, v = (torch.randn(2, 4, n, 16, generator=g, dtype=F64) for n in (7, 9, 9))
    mask = torch.zeros(2, 1, 1, 9, dtype=torch.float32); mask[0, ..., 6:] = 1; mask[1, ..., 3:] = 1
    o, w = utils.scaled_dot_product_attention(q, k, v, mask)
    o2, w2 = O.scaled_dot_product_attention(q, k, v, mask.to(F64))
    assert torch.allclose(o, o2, atol=1e-13) and torch.allclose(w, w2, atol=1e-13)


def test_warmup_decay_schedule(ref):
    _, sched, _ = ref
    for d in (128, 256):
        s = sched.WarmupDecay(d, warmup_steps=5000)                  # models/sketchformer.py:113-114
        for step in (1, 7, 4999, 5000, 5001, 20000, 123456):
            got = float(s(torch.tensor(float(step), dtype=torch.float32)))
            assert got == pytest.approx(O.warmup_decay_lr(step, d), rel=2e-6), (d, step)
    assert math.isclose(O.warmup_decay_lr(5000, 256), 256 ** -0.5 * 5000 ** -0.5)


def test_encoder_stack(ref):
    utils, _, tr = ref
    cfg = O.OracleConfig(**TINY)
    p = _randomise(O.init_params(cfg, 3, F64), 4)
    x = _batch(cfg)
    enc = tr.Encoder(cfg.num_layers, cfg.d_model, cfg.num_heads, cfg.dff, cfg.vocab_size, rate=0.1)
    enc.embedding.embeddings = p["enc.emb"].clone()
    for i, lay in enumerate(enc.enc_layers):
        _set_mha(lay.mha, p, f"enc.l{i}.mha")
        _set_ffn(lay.ffn, p, f"enc.l{i}.ffn")
        _set_ln(lay.layernorm1, p, f"enc.l{i}.ln1")
        _set_ln(lay.layernorm2, p, f"enc.l{i}.ln2")
    mask = utils.create_padding_mask(x)
    got = enc(x, False, mask)
    want = O.encoder(p, x, O.create_padding_mask(x, F64), cfg, False, None, None, F64)
    assert got.dtype == F64 and torch.allclose(got, want, atol=1e-12), float((got - want).abs().max())


@pytest.mark.parametrize("blind", [True, False])
def test_decoder_stack(ref, blind):
    utils, _, tr = ref
    cfg = O.OracleConfig(**TINY)
    p = _randomise(O.init_params(cfg, 5, F64), 6)
    x = _batch(cfg)
    tar = x[:, :-1]
    dec = tr.Decoder(cfg.num_layers, cfg.d_model, cfg.num_heads, cfg.dff, cfg.vocab_size, rate=0.1)
    dec.embedding.embeddings = p["dec.emb"].clone()
    for i, lay in enumerate(dec.dec_layers):
        _set_mha(lay.mha1, p, f"dec.l{i}.mha1")
        _set_mha(lay.mha2, p, f"dec.l{i}.mha2")
        _set_ffn(lay.ffn, p, f"dec.l{i}.ffn")
        for j in (1, 2, 3):
            _set_ln(getattr(lay, f"layernorm{j}"), p, f"dec.l{i}.ln{j}")
    g = torch.Generator().manual_seed(9)
    memory = torch.randn(x.shape[0], cfg.max_seq_len, cfg.d_model, generator=g, dtype=F64)
    _, combined, dec_pad = utils.create_masks(x, tar)
    pad = torch.zeros_like(dec_pad) if blind else dec_pad          # models/sketchformer.py:172-173
    got, attn = dec(tar, memory, False, combined, pad)
    _, comb_o, pad_o = O.create_masks(x, tar, F64)
    want, attn_o = O.decoder(p, tar, memory, comb_o, torch.zeros_like(pad_o) if blind else pad_o, cfg, False, None,
                             None, F64)
    assert torch.allclose(got, want, atol=1e-12), float((got - want).abs().max())
    assert set(attn) == set(attn_o)
    for k in attn:
        assert torch.allclose(attn[k], attn_o[k], atol=1e-12)


@pytest.mark.parametrize("version", [1, 2])
def test_pooling_bottleneck_and_expander(ref, version):
    _, _, tr = ref
    cfg = O.OracleConfig(**dict(TINY, attn_version=version))
    p = _randomise(O.init_params(cfg, 7, F64), 8)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(4, cfg.max_seq_len, cfg.d_model, generator=g, dtype=F64)
    pool = (tr.SelfAttnV1 if version == 1 else tr.SelfAttnV2)(cfg.lowerdim)      # models/sketchformer.py:96-98
    pool(x)                                                       # builds W_attn / b_attn / V_attn (and the Dense)
    assert tuple(pool.W.shape) == tuple(p["pool.W"].shape) and tuple(pool.V.shape) == tuple(p["pool.V"].shape)
    pool.W, pool.b, pool.V = p["pool.W"].clone(), p["pool.b"].clone(), p["pool.V"].clone()
    if version != 1:
        _set_dense(pool.embeding_layer, p, "pool.emb")
    emb, a = pool(x)
    emb_o, a_o = O.bottleneck(p, x, cfg)
    assert torch.allclose(emb, emb_o, atol=1e-13) and torch.allclose(a, a_o, atol=1e-13)
    assert emb.shape == (4, cfg.d_model if version == 1 else cfg.lowerdim)
    exp = tr.DenseExpander(cfg.max_seq_len)
    exp(emb)
    assert tuple(exp.expand_layer.kernel.shape) == tuple(p["exp.k"].shape)
    _set_dense(exp.expand_layer, p, "exp")
    pre = exp(emb)
    assert pre.shape == (4, cfg.max_seq_len, emb.shape[1])
    assert torch.allclose(pre, O.dense_expander(p, emb_o), atol=1e-13)


def test_multi_head_attention_argument_order(ref):
    """MultiHeadAttention.call takes (v, k, q, mask) (transformer.py:167); cross attention with different key and
    query lengths tells the three apart."""
    _, _, tr = ref
    cfg = O.OracleConfig(**TINY)
    p = _randomise(O.init_params(cfg, 11, F64), 12)
    g = torch.Generator().manual_seed(3)
    q = torch.randn(3, 5, cfg.d_model, generator=g, dtype=F64)
    k = torch.randn(3, 9, cfg.d_model, generator=g, dtype=F64)
    v = torch.randn(3, 9, cfg.d_model, generator=g, dtype=F64)
    m = tr.MultiHeadAttention(cfg.d_model, cfg.num_heads)
    _set_mha(m, p, "dec.l0.mha2")
    mask = torch.zeros(3, 1, 1, 9, dtype=torch.float32); mask[1, ..., 4:] = 1
    got, w = m(v, k, q, mask)
    want, w_o = O.mha(p, "dec.l0.mha2", v, k, q, mask.to(F64), cfg.num_heads)
    assert got.shape == (3, 5, cfg.d_model)
    assert torch.allclose(got, want, atol=1e-13) and torch.allclose(w, w_o, atol=1e-13)

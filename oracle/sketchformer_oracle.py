"""CPU oracle: an op-for-op restatement of the reference's TF-2.1 graph in torch.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (``sketchformer_b200``)
may import this module; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and there only
as the checker or as the reported CPU baseline.

HOW FAR PARITY IS PINNED.  The reference (leosampaio/sketchformer) ships no tests and no golden
vectors, and TensorFlow 2.1 cannot be installed here (no network), so the reference cannot run as
shipped and NO TensorFlow-computed number pins this file: against TensorFlow's own kernels parity
stays UNPINNED.  What is pinned is everything the reference itself wrote in Python: its
builders/utils.py, builders/layers/transformer.py, builders/losses.py, builders/schedulers.py and
models/sketchformer.py are EXECUTED from /root/reference on a torch-backed stand-in for the few
dozen TF primitives they call (``oracle/tf_shim.py``; Dense / Embedding / LayerNormalization /
softmax / matmul / cast ... restated from the public API documentation), with this oracle's
parameters loaded into the reference's layer objects, and must agree with the functions below to
1e-12 in float64: forward outputs, both losses, the schedule, the optimizer hyper-parameters that
``build_model`` passes, class prediction and greedy decode, for six graph-changing hparam variants
(``tests/test_reference_layers.py``, ``tests/test_reference_model.py``; vectors frozen by
``oracle/make_reference_golden.py`` in ``tests/golden/reference_model_tiny.npz`` for boxes
without /root/reference).  In addition: fp64 vs fp32, finite differences,
``torch.nn.functional`` and ``torch.nn.Transformer*Layer`` cross-checks (``tests/test_oracle.py``).

Every function cites the reference file:line it restates (paths relative to the
reference repository root).  TF semantics that are not visible in the reference
sources (Keras initialisers, LayerNormalization, Adam, losses) follow TF 2.1.

Parameter naming (flat dict ``name -> tensor``; Keras shapes, kernel = (in, out)):
    enc.emb (V,d) | enc.emb.k (5,d), enc.emb.b (d)       [continuous]
    enc.l{i}.mha.{wq,wk,wv,wo}.{k,b}; enc.l{i}.ffn.{w1,w2}.{k,b}; enc.l{i}.ln{1,2}.{g,b}
    pool.W (d,u) pool.b (u) pool.V (u,1) [+ pool.emb.{k,b} for attn_version 2]
    cbuf{j}.{k,b}; cls.{k,b}; exp.k (1,T) exp.b (T)
    dec.emb ...; dec.l{i}.{mha1,mha2}.{wq,wk,wv,wo}.{k,b}; dec.l{i}.ffn...; dec.l{i}.ln{1,2,3}.{g,b}
    out.{k,b}
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field, asdict
from typing import Dict, Optional

import numpy as np
import torch

Tensor = torch.Tensor


@dataclass
class OracleConfig:
    # models/sketchformer.py:24-53 (model hparams) + dataloaders/distributed_stroke3.py:15-25
    num_layers: int = 4
    d_model: int = 128
    dff: int = 512
    num_heads: int = 8
    dropout_rate: float = 0.0          # parity runs use 0
    lowerdim: int = 256
    attn_version: int = 1
    do_classification: bool = True
    class_weight: float = 1.0
    class_buffer_layers: int = 0
    class_dropout: float = 0.1         # models/sketchformer.py:38
    do_reconstruction: bool = True
    recon_weight: float = 1.0
    blind_decoder_mask: bool = True
    max_seq_len: int = 200
    use_continuous_data: bool = False
    vocab_size: int = 1004             # utils/tokenizer.py:29-33 with n_clusters=1000
    n_classes: int = 345
    warmup_steps: int = 5000           # hard-coded at models/sketchformer.py:113-114
    max_pos: int = 1000                # builders/layers/transformer.py:268,307

    def to_dict(self):
        return asdict(self)


# ---------------------------------------------------------------------------
# builders/utils.py
# ---------------------------------------------------------------------------

def positional_encoding(position: int, d_model: int) -> Tensor:
    """builders/utils.py:17-32 — float64 numpy, interleaved sin/cos, cast fp32."""
    pos = np.arange(position)[:, np.newaxis]
    i = np.arange(d_model)[np.newaxis, :]
    angle_rates = 1 / np.power(10000, (2 * (i // 2)) / np.float32(d_model))
    angle_rads = pos * angle_rates
    angle_rads[:, 0::2] = np.sin(angle_rads[:, 0::2])
    angle_rads[:, 1::2] = np.cos(angle_rads[:, 1::2])
    return torch.from_numpy(angle_rads[np.newaxis, ...].astype(np.float32))


def create_padding_mask(seq: Tensor, dtype) -> Tensor:
    """builders/utils.py:35-43 — 1.0 where PAD (token==0, or last stroke-5 bit==1)."""
    if seq.dim() < 3:
        m = (seq == 0).to(dtype)
    else:
        m = (seq[..., -1] == 1).to(dtype)
    return m[:, None, None, :]


def create_look_ahead_mask(size: int, dtype) -> Tensor:
    """builders/utils.py:46-49 — strict upper triangle of ones."""
    return 1 - torch.tril(torch.ones(size, size, dtype=dtype))


def create_masks(inp: Tensor, tar: Tensor, dtype):
    """builders/utils.py:52-68."""
    enc_padding_mask = create_padding_mask(inp, dtype)
    dec_padding_mask = create_padding_mask(inp, dtype)
    look_ahead = create_look_ahead_mask(tar.shape[1], dtype)
    dec_target_padding_mask = create_padding_mask(tar, dtype)
    combined = torch.maximum(dec_target_padding_mask, look_ahead)
    return enc_padding_mask, combined, dec_padding_mask


def scaled_dot_product_attention(q, k, v, mask):
    """builders/utils.py:71-105 — matmul, /sqrt(dk), += mask*-1e9, softmax, matmul."""
    matmul_qk = torch.matmul(q, k.transpose(-1, -2))
    dk = torch.tensor(float(k.shape[-1]), dtype=q.dtype)
    logits = matmul_qk / torch.sqrt(dk)
    if mask is not None:
        logits = logits + (mask * -1e9)
    w = torch.softmax(logits, dim=-1)
    return torch.matmul(w, v), w


# ---------------------------------------------------------------------------
# Parameters (Keras default initialisers)
# ---------------------------------------------------------------------------

def _glorot(gen, fan_in, fan_out, shape):
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * limit


def _uniform05(gen, shape):
    return (torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * 0.05


def _dense(p, name, gen, fin, fout):
    p[name + ".k"] = _glorot(gen, fin, fout, (fin, fout))
    p[name + ".b"] = torch.zeros(fout, dtype=torch.float64)


def _mha(p, name, gen, d, kv_in):
    _dense(p, name + ".wq", gen, d, d)
    _dense(p, name + ".wk", gen, kv_in, d)
    _dense(p, name + ".wv", gen, kv_in, d)
    _dense(p, name + ".wo", gen, d, d)


def _ln(p, name, d):
    p[name + ".g"] = torch.ones(d, dtype=torch.float64)
    p[name + ".b"] = torch.zeros(d, dtype=torch.float64)


def init_params(cfg: OracleConfig, seed: int = 0, dtype=torch.float32) -> Dict[str, Tensor]:
    """Create every trainable variable build_model() would (models/sketchformer.py:63-107),
    in a fixed documented order, from one torch.Generator."""
    gen = torch.Generator().manual_seed(seed)
    d, f, L, T = cfg.d_model, cfg.dff, cfg.num_layers, cfg.max_seq_len
    p: Dict[str, Tensor] = {}

    def embed(prefix):
        if cfg.use_continuous_data:     # transformer.py:275-276
            _dense(p, prefix + ".emb", gen, 5, d)
        else:                            # transformer.py:278 (RandomUniform +-0.05)
            p[prefix + ".emb"] = _uniform05(gen, (cfg.vocab_size, d))

    embed("enc")
    for i in range(L):                   # transformer.py:201-213
        _mha(p, f"enc.l{i}.mha", gen, d, d)
        _dense(p, f"enc.l{i}.ffn.w1", gen, d, f)
        _dense(p, f"enc.l{i}.ffn.w2", gen, f, d)
        _ln(p, f"enc.l{i}.ln1", d)
        _ln(p, f"enc.l{i}.ln2", d)
    emb_dim = d
    if cfg.lowerdim:
        u = cfg.lowerdim
        if cfg.attn_version == 1:        # SelfAttnV1.build transformer.py:41-59
            p["pool.W"] = torch.randn((d, u), generator=gen, dtype=torch.float64) * 0.05
            p["pool.b"] = torch.zeros(u, dtype=torch.float64)
            p["pool.V"] = _uniform05(gen, (u, 1))
        else:                            # SelfAttnV2.build transformer.py:96-114
            p["pool.W"] = torch.randn((d, d), generator=gen, dtype=torch.float64) * 0.05
            p["pool.b"] = torch.zeros(d, dtype=torch.float64)
            p["pool.V"] = _uniform05(gen, (d, 1))
            _dense(p, "pool.emb", gen, d, u)
            emb_dim = u
        if cfg.do_classification:        # sketchformer.py:99-104
            fin = emb_dim
            for j in range(cfg.class_buffer_layers):
                _dense(p, f"cbuf{j}", gen, fin, u)
                fin = u
            _dense(p, "cls", gen, fin, cfg.n_classes)
        if cfg.do_reconstruction:        # DenseExpander transformer.py:359-364
            p["exp.k"] = _glorot(gen, 1, T, (1, T))
            p["exp.b"] = torch.zeros(T, dtype=torch.float64)
    if cfg.do_reconstruction:
        embed("dec")
        for i in range(L):               # transformer.py:227-243
            _mha(p, f"dec.l{i}.mha1", gen, d, d)
            _mha(p, f"dec.l{i}.mha2", gen, d, emb_dim)
            _dense(p, f"dec.l{i}.ffn.w1", gen, d, f)
            _dense(p, f"dec.l{i}.ffn.w2", gen, f, d)
            _ln(p, f"dec.l{i}.ln1", d)
            _ln(p, f"dec.l{i}.ln2", d)
            _ln(p, f"dec.l{i}.ln3", d)
        nout = 5 if cfg.use_continuous_data else cfg.vocab_size   # sketchformer.py:86,91
        _dense(p, "out", gen, d, nout)
    return {k: v.to(dtype) for k, v in p.items()}


def count_params(p: Dict[str, Tensor]) -> int:
    return sum(int(v.numel()) for v in p.values())


# ---------------------------------------------------------------------------
# Layers (builders/layers/transformer.py)
# ---------------------------------------------------------------------------

def dense(p, name, x):
    return torch.matmul(x, p[name + ".k"]) + p[name + ".b"]


def layer_norm(p, name, x, eps=1e-6):
    """tf.keras.layers.LayerNormalization(epsilon=1e-6): biased variance over last axis."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + eps) * p[name + ".g"] + p[name + ".b"]


_DROPOUT_MASKS = None   # test hook, see dropout_masks()


class dropout_masks:
    """Context manager for the parity tests: inside it every Dropout layer takes its keep-mask from
    `provider(site_name, shape) -> bool array` instead of torch's RNG (TF's stateful dropout RNG cannot be
    reproduced by any other implementation; the CUDA path's counter-based masks can, oracle/dropout_rng.py).
    Site names: enc.x0, enc.l{i}.attn, enc.l{i}.ffn, dec.x0, dec.l{i}.attn1, dec.l{i}.attn2, dec.l{i}.ffn."""

    def __init__(self, provider):
        self.provider = provider

    def __enter__(self):
        global _DROPOUT_MASKS
        self.prev, _DROPOUT_MASKS = _DROPOUT_MASKS, self.provider
        return self

    def __exit__(self, *exc):
        global _DROPOUT_MASKS
        _DROPOUT_MASKS = self.prev


def _dropout(x, rate, training, gen, site=None):
    """tf.keras.layers.Dropout: inverted dropout, x * keep * 1/(1-rate) when training."""
    if not training or rate == 0.0:
        return x
    if _DROPOUT_MASKS is not None and site is not None:
        keep = torch.as_tensor(_DROPOUT_MASKS(site, tuple(x.shape))).to(x.dtype)
    else:
        keep = (torch.rand(x.shape, generator=gen) >= rate).to(x.dtype)
    scale = torch.tensor(1.0, dtype=torch.float32) / (torch.tensor(1.0, dtype=torch.float32) - torch.tensor(rate, dtype=torch.float32))
    return x * keep * scale.to(x.dtype)


def split_heads(x, H):
    """transformer.py:160-165."""
    B, T, d = x.shape
    return x.reshape(B, T, H, d // H).permute(0, 2, 1, 3)


def mha(p, name, v, k, q, mask, H, tape=None):
    """MultiHeadAttention.call transformer.py:167-191 (argument order v, k, q)."""
    B = q.shape[0]
    qq = dense(p, name + ".wq", q)
    kk = dense(p, name + ".wk", k)
    vv = dense(p, name + ".wv", v)
    qh, kh, vh = split_heads(qq, H), split_heads(kk, H), split_heads(vv, H)
    att, w = scaled_dot_product_attention(qh, kh, vh, mask)
    att = att.permute(0, 2, 1, 3)
    concat = att.reshape(B, -1, qq.shape[-1])
    if tape is not None:
        tape[name + ".q"] = qq
        tape[name + ".k"] = kk
        tape[name + ".v"] = vv
        tape[name + ".ctx"] = concat
    return dense(p, name + ".wo", concat), w


_RELU_PATTERN = None   # test hook, see relu_pattern()


class relu_pattern:
    """Context manager for the parity tests: inside it, ffn() uses the given 0/1 activation pattern
    (name -> bool tensor shaped like the hidden layer) instead of recomputing `pre > 0`.  ReLU is not
    differentiable at 0: a forward perturbation eps flips a fraction ~eps of the units and moves the
    gradient by ~sqrt(eps) in L2, so gradients are compared at the SAME activation pattern and the
    pattern itself is compared separately (tests/test_gpu_train.py)."""

    def __init__(self, pattern=None, record=None):
        self.pattern = dict(pattern or {})
        self.pattern["__record__"] = record      # dict filled with the oracle's own pattern, or None

    def __enter__(self):
        global _RELU_PATTERN
        self.prev, _RELU_PATTERN = _RELU_PATTERN, self.pattern
        return self

    def __exit__(self, *exc):
        global _RELU_PATTERN
        _RELU_PATTERN = self.prev


def ffn(p, name, x):
    """point_wise_feed_forward_network transformer.py:194-198."""
    pre = dense(p, name + ".w1", x)
    pat = _RELU_PATTERN.get(name) if _RELU_PATTERN else None
    if _RELU_PATTERN and _RELU_PATTERN.get("__record__") is not None:
        _RELU_PATTERN["__record__"][name] = (pre > 0).detach()
    h = torch.relu(pre) if pat is None else pre * pat.to(pre.dtype)
    return dense(p, name + ".w2", h)


def embed_input(p, prefix, x, cfg, dtype):
    """Encoder.call / Decoder.call head: transformer.py:288-296, 325-334."""
    T = x.shape[1]
    if cfg.use_continuous_data:
        h = dense(p, prefix + ".emb", x.to(dtype))
    else:
        h = p[prefix + ".emb"][x]
    h = h * torch.sqrt(torch.tensor(float(cfg.d_model), dtype=torch.float32)).to(dtype)
    h = h + positional_encoding(cfg.max_pos, cfg.d_model)[:, :T, :].to(dtype)
    return h


def encoder(p, x, mask, cfg, training, gen, tape, dtype):
    """Encoder.call + EncoderLayer.call transformer.py:215-224, 288-301."""
    h = embed_input(p, "enc", x, cfg, dtype)
    if tape is not None:
        tape["enc.x0"] = h
    h = _dropout(h, cfg.dropout_rate, training, gen, "enc.x0")
    for i in range(cfg.num_layers):
        n = f"enc.l{i}"
        a, _ = mha(p, n + ".mha", h, h, h, mask, cfg.num_heads, tape)
        a = _dropout(a, cfg.dropout_rate, training, gen, n + ".attn")
        out1 = layer_norm(p, n + ".ln1", h + a)
        ff = ffn(p, n + ".ffn", out1)
        ff = _dropout(ff, cfg.dropout_rate, training, gen, n + ".ffn")
        h = layer_norm(p, n + ".ln2", out1 + ff)
        if tape is not None:
            tape[n + ".out1"] = out1
            tape[n + ".out2"] = h
    return h


def bottleneck(p, x, cfg, tape=None):
    """SelfAttnV1.call transformer.py:61-74 / SelfAttnV2.call 116-131.  No padding mask."""
    ui = torch.tanh(torch.matmul(x, p["pool.W"]) + p["pool.b"])
    ai = torch.softmax(torch.matmul(ui, p["pool.V"]), dim=1)      # (B,T,1)
    o = (x * ai).sum(dim=1)
    if cfg.attn_version != 1:
        o = dense(p, "pool.emb", o)
    if tape is not None:
        tape["pool.a"] = ai[..., 0]
    return o, ai


def classify_from_embedding(p, emb, cfg, training, gen):
    """models/sketchformer.py:183-199; returns (probabilities, logits)."""
    fc = emb
    for j in range(cfg.class_buffer_layers):
        fc = torch.relu(dense(p, f"cbuf{j}", fc))
        fc = _dropout(fc, cfg.class_dropout, training, gen, f"cbuf{j}")
    logits = dense(p, "cls", fc)
    return torch.softmax(logits, dim=-1), logits


def dense_expander(p, emb):
    """DenseExpander.call transformer.py:370-376: Dense(seq_len) on a trailing
    singleton axis then transpose -> pre[b,t,c] = emb[b,c]*w[t] + beta[t]."""
    x = emb.unsqueeze(2)                                          # (B,d,1)
    x = torch.matmul(x, p["exp.k"]) + p["exp.b"]                  # (B,d,T)
    return x.permute(0, 2, 1)


def decoder(p, x, enc_output, look_ahead_mask, padding_mask, cfg, training, gen, tape, dtype):
    """Decoder.call + DecoderLayer.call transformer.py:245-262, 325-344."""
    attention_weights = {}
    h = embed_input(p, "dec", x, cfg, dtype)
    if tape is not None:
        tape["dec.x0"] = h
    h = _dropout(h, cfg.dropout_rate, training, gen, "dec.x0")
    for i in range(cfg.num_layers):
        n = f"dec.l{i}"
        a1, w1 = mha(p, n + ".mha1", h, h, h, look_ahead_mask, cfg.num_heads, tape)
        a1 = _dropout(a1, cfg.dropout_rate, training, gen, n + ".attn1")
        out1 = layer_norm(p, n + ".ln1", a1 + h)
        a2, w2 = mha(p, n + ".mha2", enc_output, enc_output, out1, padding_mask, cfg.num_heads, tape)
        a2 = _dropout(a2, cfg.dropout_rate, training, gen, n + ".attn2")
        out2 = layer_norm(p, n + ".ln2", a2 + out1)
        ff = ffn(p, n + ".ffn", out2)
        ff = _dropout(ff, cfg.dropout_rate, training, gen, n + ".ffn")
        h = layer_norm(p, n + ".ln3", ff + out2)
        attention_weights[f"decoder_layer{i + 1}_block1"] = w1
        attention_weights[f"decoder_layer{i + 1}_block2"] = w2
        if tape is not None:
            tape[n + ".out1"] = out1
            tape[n + ".out2"] = out2
            tape[n + ".out3"] = h
    return h, attention_weights


# ---------------------------------------------------------------------------
# Model (models/sketchformer.py)
# ---------------------------------------------------------------------------

def encode(p, inp, inp_mask, cfg, training=False, gen=None, tape=None, dtype=torch.float32):
    """Transformer.encode models/sketchformer.py:149-160."""
    out = {"enc_output": encoder(p, inp, inp_mask, cfg, training, gen, tape, dtype), "class": None}
    if cfg.lowerdim:
        emb, _ = bottleneck(p, out["enc_output"], cfg, tape)
        out["embedding"] = emb
        if cfg.do_classification:
            out["class"], out["class_logits"] = classify_from_embedding(p, emb, cfg, training, gen)
    else:
        out["embedding"] = out["enc_output"]
    return out


def decode(p, embedding, target, target_mask, look_ahead_mask, cfg, training=False, gen=None,
           tape=None, dtype=torch.float32):
    """Transformer.decode models/sketchformer.py:170-181."""
    padding_mask = torch.zeros_like(target_mask) if cfg.blind_decoder_mask else target_mask
    pre = dense_expander(p, embedding) if cfg.lowerdim else embedding
    if tape is not None:
        tape["pre_decoder"] = pre
    dec_out, attn = decoder(p, target, pre, look_ahead_mask, padding_mask, cfg, training, gen, tape, dtype)
    final = dense(p, "out", dec_out)
    return {"recon": final, "attn_weights": attn}


def call(p, inp, tar, cfg, training, masks, gen=None, tape=None, dtype=torch.float32):
    """Transformer.call models/sketchformer.py:131-147."""
    enc_pad, look_ahead, dec_pad = masks
    enc = encode(p, inp, enc_pad, cfg, training, gen, tape, dtype)
    out = {k: enc[k] for k in ("embedding", "class", "class_logits") if k in enc}
    if cfg.do_reconstruction:
        dec = decode(p, enc["embedding"], tar, dec_pad, look_ahead, cfg, training, gen, tape, dtype)
        out["recon"] = dec["recon"]
    return out


# ---------------------------------------------------------------------------
# Losses (builders/losses.py) and metrics (builders/keras_metrics.py)
# ---------------------------------------------------------------------------

def class_loss(labels, probs_logits):
    """builders/losses.py:21-24 — Keras SCCE on a Softmax output; in graph mode TF-2.1
    back-tracks to the logits, i.e. -log_softmax(z)[y], batch mean."""
    logits = probs_logits
    lsm = torch.log_softmax(logits, dim=-1)
    y = labels.reshape(-1).long()
    return -(lsm.gather(1, y[:, None])[:, 0]).mean()


def reconstruction_loss(real, pred):
    """builders/losses.py:26-41 — mean over ALL B*Td of CE*(real!=0)."""
    mask = (real != 0).to(pred.dtype)
    lse = torch.logsumexp(pred, dim=-1)
    tgt = pred.gather(-1, real.long().unsqueeze(-1))[..., 0]
    loss_ = (lse - tgt) * mask
    return loss_.mean()


def continuous_reconstruction_loss(real, pred):
    """builders/losses.py:43-66."""
    real = real.to(pred.dtype)
    mask = (real[..., -1] != 1).to(pred.dtype)
    location_loss = ((real[:, :, :2] - pred[:, :, :2]) ** 2).mean(dim=-1)
    labels = real[:, :, 2:].argmax(dim=-1)
    lsm = torch.log_softmax(pred[:, :, 2:], dim=-1)
    metadata_loss = -(lsm.gather(-1, labels.unsqueeze(-1))[..., 0])
    metadata_loss = metadata_loss.mean()
    loss_ = (location_loss + metadata_loss) * mask
    return loss_.mean()


def sparse_categorical_accuracy(y_true, y_pred):
    """tf.keras.metrics.SparseCategoricalAccuracy: mean(argmax(pred)==true) over all elements."""
    return (y_pred.argmax(dim=-1) == y_true.reshape(y_pred.shape[:-1]).long()).to(torch.float64).mean()


def compute_losses(out, tar_real, lab, cfg):
    """Loss section of model_trainer, models/sketchformer.py:332-346."""
    res = {}
    total = 0.0
    if cfg.do_reconstruction:
        if cfg.use_continuous_data:
            recon = cfg.recon_weight * continuous_reconstruction_loss(tar_real, out["recon"])
        else:
            recon = cfg.recon_weight * reconstruction_loss(tar_real, out["recon"])
            res["recon_acc"] = sparse_categorical_accuracy(tar_real, out["recon"])
        res["recon_loss"] = recon
        total = total + recon
    if cfg.do_classification and cfg.lowerdim:
        clas = cfg.class_weight * class_loss(lab, out["class_logits"])
        res["class_loss"] = clas
        res["class_acc"] = sparse_categorical_accuracy(lab, out["class"])
        total = total + clas
    res["total_loss"] = total
    return res


# ---------------------------------------------------------------------------
# Schedule + optimiser (builders/schedulers.py:13-29; tf.keras.optimizers.Adam)
# ---------------------------------------------------------------------------

def warmup_decay_lr(step: int, d_model: int, warmup_steps: int = 5000) -> float:
    """WarmupDecay.__call__: d^-0.5 * min(step^-0.5, step*warmup^-1.5); lr(0) = 0."""
    step = float(step)
    arg1 = float("inf") if step == 0 else step ** -0.5
    arg2 = step * (warmup_steps ** -1.5)
    return (float(d_model) ** -0.5) * min(arg1, arg2)


class AdamTF:
    """TF-2.1 Adam(beta_1=0.9, beta_2=0.98, epsilon=1e-9), models/sketchformer.py:121-123;
    update rule of training_ops ApplyAdam: epsilon outside bias correction."""

    def __init__(self, params, d_model, beta1=0.9, beta2=0.98, eps=1e-9, warmup_steps=5000):
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}
        self.iterations = 0
        self.b1, self.b2, self.eps = beta1, beta2, eps
        self.d_model, self.warmup = d_model, warmup_steps

    def apply(self, params, grads):
        lr = warmup_decay_lr(self.iterations, self.d_model, self.warmup)
        t = self.iterations + 1
        alpha = lr * math.sqrt(1.0 - self.b2 ** t) / (1.0 - self.b1 ** t)
        with torch.no_grad():
            for k, w in params.items():
                g = grads[k]
                self.m[k] += (g - self.m[k]) * (1.0 - self.b1)
                self.v[k] += (g * g - self.v[k]) * (1.0 - self.b2)
                w -= (self.m[k] * alpha) / (torch.sqrt(self.v[k]) + self.eps)
        self.iterations += 1
        return lr


# ---------------------------------------------------------------------------
# Train step (models/sketchformer.py:325-349) and inference (201-311)
# ---------------------------------------------------------------------------

def forward_losses(p, inp, lab, cfg, training=True, gen=None, tape=None, dtype=torch.float32):
    """inp = tar = data (sketchformer.py:354); tar_inp = tar[:, :-1]; tar_real = tar[:, 1:]."""
    tar_inp = inp[:, :-1, ...]
    tar_real = inp[:, 1:, ...]
    masks = create_masks(inp, tar_inp, dtype)
    out = call(p, inp, tar_inp, cfg, training, masks, gen, tape, dtype)
    losses = compute_losses(out, tar_real, lab, cfg)
    return out, losses


def train_step(p, opt: AdamTF, inp, lab, cfg, gen=None, tape=None, dtype=torch.float32):
    """One model_trainer call; returns (losses dict of floats, grads dict)."""
    for v in p.values():
        v.requires_grad_(True)
        v.grad = None
    out, losses = forward_losses(p, inp, lab, cfg, True, gen, tape, dtype)
    losses["total_loss"].backward()
    grads = {k: (v.grad.detach().clone() if v.grad is not None else torch.zeros_like(v))
             for k, v in p.items()}
    for v in p.values():
        v.requires_grad_(False)
        v.grad = None
    opt.apply(p, grads)
    return {k: float(v.detach()) if torch.is_tensor(v) else float(v) for k, v in losses.items()}, grads


class RunningMetrics:
    """builders/keras_metrics.py:13-42 — Mean / SparseCategoricalAccuracy never reset in train()."""

    def __init__(self):
        self.sum: Dict[str, float] = {}
        self.cnt: Dict[str, float] = {}

    def update(self, losses: Dict[str, float], weights: Optional[Dict[str, float]] = None):
        for k, v in losses.items():
            w = (weights or {}).get(k, 1.0)
            self.sum[k] = self.sum.get(k, 0.0) + float(v) * w
            self.cnt[k] = self.cnt.get(k, 0.0) + w

    def result(self):
        return {k: self.sum[k] / self.cnt[k] for k in self.sum}


def encode_from_seq(p, inp_seq, cfg, dtype=torch.float32):
    """models/sketchformer.py:162-168."""
    mask = create_padding_mask(inp_seq, dtype)
    return encode(p, inp_seq, mask, cfg, False, None, None, dtype)


def predict_class(p, inp_seq, cfg, dtype=torch.float32):
    """models/sketchformer.py:223-228."""
    out = encode_from_seq(p, inp_seq, cfg, dtype)
    if cfg.do_classification:
        out["class"] = out["class"].argmax(dim=-1).to(torch.int32)
    return out


def make_dummy_input(cfg, expected_len, nattn, batch_size, dtype):
    """models/sketchformer.py:230-253."""
    T = cfg.max_seq_len
    if cfg.use_continuous_data:
        d = torch.zeros(batch_size, T, 5, dtype=dtype)
        d[:, nattn:, 4] = 1.0
        return d
    if expected_len is None:
        d = torch.zeros(batch_size, T, dtype=dtype)
        d[:, :nattn] = 1
        return d
    d = torch.zeros(batch_size, T, dtype=dtype)
    for b in range(batch_size):
        d[b, :int(nattn[b])] = 1
    return d


def predict_from_embedding(p, emb, cfg, expected_len=None, sos=1002, eos=1003, dtype=torch.float32):
    """Greedy decode, models/sketchformer.py:255-311: re-runs the decoder on the growing
    prefix; stops when every row has emitted EOS (tokens) / pen state 2 (continuous)."""
    out = {}
    B = emb.shape[0]
    embedding = emb.to(dtype)
    if cfg.do_reconstruction:
        if cfg.use_continuous_data:
            output = torch.ones(B, 1, 5, dtype=dtype) * torch.tensor([0., 0., 1., 0., 0.], dtype=dtype)
        else:
            output = torch.full((B, 1), sos, dtype=torch.int32)
            eos_lst = [0] * B
        res = None
        for i in range(cfg.max_seq_len):
            nattn = expected_len if expected_len is not None else i + 1
            dummy = make_dummy_input(cfg, expected_len, nattn, B, dtype)
            _, combined, dec_pad = create_masks(dummy, output, dtype)
            res = decode(p, embedding, output if cfg.use_continuous_data else output.long(),
                         dec_pad, combined, cfg, False, None, None, dtype)
            predictions = res["recon"][:, -1:, ...]
            if cfg.use_continuous_data:
                predicted = torch.cat((predictions[..., :2],
                                       torch.softmax(predictions[..., 2:], dim=-1)), dim=-1)
                finished = int((predicted[..., 2:].argmax(dim=-1) == 2).sum())
                output = torch.cat([output, predicted], dim=1)
                if finished == B:
                    break
            else:
                predicted = predictions.argmax(dim=-1).to(torch.int32)
                output = torch.cat([output, predicted], dim=1)
                for b in torch.nonzero(predicted.reshape(-1) == eos).reshape(-1).tolist():
                    eos_lst[b] = 1
                if sum(eos_lst) == B:
                    break
        out["recon"] = output
        out["attn_weights"] = res["attn_weights"]
    if cfg.do_classification:
        probs, _ = classify_from_embedding(p, embedding, cfg, False, None)
        out["class"] = probs.argmax(dim=-1).to(torch.int32)
    return out


def predict(p, inp_seq, cfg, sos=1002, eos=1003, dtype=torch.float32):
    """models/sketchformer.py:201-221."""
    out = encode_from_seq(p, inp_seq, cfg, dtype)
    if cfg.do_classification:
        out["class"] = out["class"].argmax(dim=-1).to(torch.int32)
    if cfg.do_reconstruction:
        if cfg.blind_decoder_mask:
            tlen = None
        elif cfg.use_continuous_data:
            tlen = (inp_seq[..., -1] != 1).to(torch.float32).sum(dim=-1)
        else:
            tlen = (inp_seq > 0).to(torch.float32).sum(dim=-1)
        dec = predict_from_embedding(p, out["embedding"], cfg, tlen, sos, eos, dtype)
        out["recon"] = dec["recon"]
        out["attn_weights"] = dec["attn_weights"]
    return out

"""Helpers that put the oracle's parameters into the reference's own layer / model objects (executed from
/root/reference on oracle/tf_shim.py).  Test infrastructure: used by tests/test_reference_*.py and by
oracle/make_reference_golden.py."""
from types import SimpleNamespace

import torch

from oracle import tf_shim  # noqa: F401

F64 = torch.float64


def randomise(p, seed):
    g = torch.Generator().manual_seed(seed)
    for k in p:
        if k.endswith(".b") or k.endswith(".g"):          # Keras leaves these at 0 / 1: make every term visible
            p[k] = p[k] + 0.1 * torch.randn(p[k].shape, generator=g, dtype=p[k].dtype)
    return p


def set_dense(layer, p, name):
    layer.kernel, layer.bias, layer.built = p[name + ".k"].clone(), p[name + ".b"].clone(), True


def set_mha(m, p, name):
    for ref_name, ours in (("wq", "wq"), ("wk", "wk"), ("wv", "wv"), ("dense", "wo")):
        set_dense(getattr(m, ref_name), p, f"{name}.{ours}")


def set_ln(layer, p, name):
    layer.gamma, layer.beta, layer.built = p[name + ".g"].clone(), p[name + ".b"].clone(), True


def set_ffn(seq, p, name):
    set_dense(seq.layers[0], p, name + ".w1")
    set_dense(seq.layers[1], p, name + ".w2")


def build_reference(ref_model, cfg, p, x, dropout_rate=0.1):
    """An instance of the reference's Transformer with the oracle's weights (BaseModel.__init__ is replaced by
    setting the attributes it would: hps, dataset; then the reference's build_model runs)."""
    Transformer, builders = ref_model
    m = object.__new__(Transformer)
    m.losses_manager = builders.losses.LossManager()                 # models/sketchformer.py:56-57
    m.metrics_manager = builders.keras_metrics.MetricManager()
    m.hps = dict(num_layers=cfg.num_layers, d_model=cfg.d_model, dff=cfg.dff, num_heads=cfg.num_heads,
                 dropout_rate=dropout_rate, lowerdim=cfg.lowerdim, attn_version=cfg.attn_version,
                 do_classification=cfg.do_classification, class_weight=cfg.class_weight,
                 class_buffer_layers=cfg.class_buffer_layers, class_dropout=cfg.class_dropout,
                 do_reconstruction=cfg.do_reconstruction, recon_weight=cfg.recon_weight,
                 blind_decoder_mask=cfg.blind_decoder_mask, optimizer="Adam", lr=0.01, lr_scheduler="WarmupDecay")
    tok = SimpleNamespace(VOCAB_SIZE=cfg.vocab_size, SOS=cfg.vocab_size - 2, EOS=cfg.vocab_size - 1, PAD=0)
    m.dataset = SimpleNamespace(hps=dict(use_continuous_data=cfg.use_continuous_data, max_seq_len=cfg.max_seq_len),
                                n_classes=cfg.n_classes, tokenizer=tok)
    m.vocab_size = None if cfg.use_continuous_data else cfg.vocab_size     # models/sketchformer.py:59-60
    m.seq_len = cfg.max_seq_len
    m.build_model()
    # one inference pass builds every lazily-built variable, then the oracle's parameters go in
    masks = m_masks(ref_model, x, x[:, :-1])
    m.call(x, x[:, :-1], False, *masks)
    if cfg.use_continuous_data:
        set_dense(m.encoder.embedding, p, "enc.emb")
    else:
        m.encoder.embedding.embeddings = p["enc.emb"].clone()
    for i, lay in enumerate(m.encoder.enc_layers):
        set_mha(lay.mha, p, f"enc.l{i}.mha"); set_ffn(lay.ffn, p, f"enc.l{i}.ffn")
        set_ln(lay.layernorm1, p, f"enc.l{i}.ln1"); set_ln(lay.layernorm2, p, f"enc.l{i}.ln2")
    if cfg.lowerdim:
        b = m.bottleneck_layer
        assert type(b).__name__ == ("SelfAttnV1" if cfg.attn_version == 1 else "SelfAttnV2")
        assert tuple(b.W.shape) == tuple(p["pool.W"].shape)
        b.W, b.b, b.V = p["pool.W"].clone(), p["pool.b"].clone(), p["pool.V"].clone()
        if cfg.attn_version != 1:
            set_dense(b.embeding_layer, p, "pool.emb")
        if cfg.do_classification:
            assert len(m.class_buffer) == cfg.class_buffer_layers
            for j, lay in enumerate(m.class_buffer):
                set_dense(lay, p, f"cbuf{j}")
            set_dense(m.classify_layer, p, "cls")
        if cfg.do_reconstruction:
            set_dense(m.expand_layer.expand_layer, p, "exp")
    else:
        assert not hasattr(m, "bottleneck_layer") and not hasattr(m, "classify_layer")
    if cfg.do_reconstruction:
        if cfg.use_continuous_data:
            set_dense(m.decoder.embedding, p, "dec.emb")
        else:
            m.decoder.embedding.embeddings = p["dec.emb"].clone()
        for i, lay in enumerate(m.decoder.dec_layers):
            set_mha(lay.mha1, p, f"dec.l{i}.mha1"); set_mha(lay.mha2, p, f"dec.l{i}.mha2")
            set_ffn(lay.ffn, p, f"dec.l{i}.ffn")
            for j in (1, 2, 3):
                set_ln(getattr(lay, f"layernorm{j}"), p, f"dec.l{i}.ln{j}")
        set_dense(m.output_layer, p, "out")
    return m


def m_masks(ref_model, inp, tar):
    return ref_model[1].utils.create_masks(inp, tar)




def trainable_variables(m):
    """Every weight tensor of a reference model built by build_reference, as torch leaves that require gradients
    (what `self.trainable_variables` of a tf.keras.Model would list); also stored on the instance."""
    out = []

    def walk(obj):
        if isinstance(obj, tf_shim.Sequential):
            for lay in obj.layers:
                walk(lay)
            return
        for name, val in list(vars(obj).items()):
            if name == "pos_encoding":
                continue
            if isinstance(val, torch.Tensor):
                val = val.detach().clone().requires_grad_(True)
                setattr(obj, name, val)
                out.append(val)
            elif isinstance(val, tf_shim.Layer):
                walk(val)
            elif isinstance(val, list) and val and all(isinstance(v, tf_shim.Layer) for v in val):
                for v in val:
                    walk(v)

    for name, val in list(vars(m).items()):
        if isinstance(val, tf_shim.Layer):
            walk(val)
        elif isinstance(val, list) and val and all(isinstance(v, tf_shim.Layer) for v in val):
            for v in val:
                walk(v)
    m.trainable_variables = out
    return out

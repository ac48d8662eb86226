"""Keras-default initialisers for every trainable variable of build_model()
(models/sketchformer.py:63-107): Dense = GlorotUniform kernel + zero bias,
Embedding = U(-0.05, 0.05), LayerNormalization = ones/zeros, SelfAttnV1 = N(0, 0.05) / zeros / U(-0.05, 0.05).
Names and shapes follow Keras (kernel = (in, out)); Engine.load_params() packs them into the arena."""
import math
from typing import Dict

import torch


def init_params(hps: dict, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    d, f, L, T = hps["d_model"], hps["dff"], hps["num_layers"], hps["max_seq_len"]
    V, C, u = hps["vocab_size"], hps["n_classes"], hps["lowerdim"]
    do_class = hps.get("do_classification", True)
    do_recon = hps.get("do_reconstruction", True)
    continuous = bool(hps.get("use_continuous_data", False))
    v2 = hps.get("attn_version", 1) != 1
    nbuf = int(hps.get("class_buffer_layers", 0))
    gen = torch.Generator().manual_seed(seed)
    p: Dict[str, torch.Tensor] = {}

    def glorot(fin, fout, shape):
        lim = math.sqrt(6.0 / (fin + fout))
        return (torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * lim

    def uni05(shape):
        return (torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * 0.05

    def dense(name, fin, fout):
        p[name + ".k"] = glorot(fin, fout, (fin, fout))
        p[name + ".b"] = torch.zeros(fout, dtype=torch.float64)

    emb_dim = u if (v2 and u) else d

    def mha(name, kv_in):
        dense(name + ".wq", d, d)
        dense(name + ".wk", kv_in, d)
        dense(name + ".wv", kv_in, d)
        dense(name + ".wo", d, d)

    def ln(name):
        p[name + ".g"] = torch.ones(d, dtype=torch.float64)
        p[name + ".b"] = torch.zeros(d, dtype=torch.float64)

    def embed(prefix):
        if continuous:                  # Dense(d_model) on stroke-5 rows, transformer.py:275-276
            dense(prefix + ".emb", 5, d)
        else:
            p[prefix + ".emb"] = uni05((V, d))

    embed("enc")
    for i in range(L):
        mha(f"enc.l{i}.mha", d)
        dense(f"enc.l{i}.ffn.w1", d, f)
        dense(f"enc.l{i}.ffn.w2", f, d)
        ln(f"enc.l{i}.ln1")
        ln(f"enc.l{i}.ln2")
    if not u:                           # lowerdim=0: no bottleneck, classifier or expander (sketchformer.py:96)
        do_class = False
    elif not v2:                        # SelfAttnV1.build transformer.py:41-59
        p["pool.W"] = torch.randn((d, u), generator=gen, dtype=torch.float64) * 0.05
        p["pool.b"] = torch.zeros(u, dtype=torch.float64)
        p["pool.V"] = uni05((u, 1))
    else:                               # SelfAttnV2.build transformer.py:96-114
        p["pool.W"] = torch.randn((d, d), generator=gen, dtype=torch.float64) * 0.05
        p["pool.b"] = torch.zeros(d, dtype=torch.float64)
        p["pool.V"] = uni05((d, 1))
        dense("pool.emb", d, u)
    if do_class:                        # models/sketchformer.py:99-104
        fin = emb_dim
        for j in range(nbuf):
            dense(f"cbuf{j}", fin, u)
            fin = u
        dense("cls", fin, C)
    if do_recon:
        if u:
            p["exp.k"] = glorot(1, T, (1, T))
            p["exp.b"] = torch.zeros(T, dtype=torch.float64)
        embed("dec")
        for i in range(L):
            mha(f"dec.l{i}.mha1", d)
            mha(f"dec.l{i}.mha2", emb_dim)
            dense(f"dec.l{i}.ffn.w1", d, f)
            dense(f"dec.l{i}.ffn.w2", f, d)
            for j in (1, 2, 3):
                ln(f"dec.l{i}.ln{j}")
        dense("out", d, 5 if continuous else V)
    return {k: v.to(dtype) for k, v in p.items()}

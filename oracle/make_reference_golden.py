"""Golden vectors made by executing the REFERENCE'S OWN model code (test infrastructure).

    python -m oracle.make_reference_golden        # needs /root/reference; writes tests/golden/reference_model_tiny.npz

models/sketchformer.py + builders/{layers/transformer,utils,losses}.py run from /root/reference on the torch-backed
TensorFlow stand-in oracle/tf_shim.py (TensorFlow itself cannot be installed here; the shim's docstring says what that
does and does not pin), in float64, with the weights of `O.init_params(cfg, seed)` perturbed by
`reference_model.randomise` and the seeded synthetic batch of `sketchformer_b200.synth`.  The fixture stores the
inputs and the reference code's outputs for every graph-changing hparam variant, so that the comparison with the
oracle also runs where /root/reference is absent (the GPU box).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import sketchformer_oracle as O                                  # noqa: E402
from oracle import tf_shim                                                    # noqa: E402
from oracle.make_golden import TINY, TINY_LOWERDIM0, TINY_NOTBLIND, TINY_V2   # noqa: E402
from oracle.reference_model import build_reference, m_masks, randomise, trainable_variables   # noqa: E402
from sketchformer_b200 import synth                                           # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "reference_model_tiny.npz")
F64 = torch.float64
VARIANTS = [("default", TINY), ("v2_classbuffer", dict(TINY_V2)), ("lowerdim0", TINY_LOWERDIM0),
            ("not_blind", TINY_NOTBLIND), ("continuous", dict(TINY, use_continuous_data=True))]
PARAM_SEED, NOISE_SEED, BATCH, BATCH_SEED = 13, 14, 4, 21


def inputs(cfg):
    if cfg.use_continuous_data:
        x, y = synth.continuous_batch(BATCH, cfg.max_seq_len, cfg.n_classes, BATCH_SEED)
        return torch.from_numpy(x).to(F64), torch.from_numpy(y)
    x, y = synth.token_batch(BATCH, cfg.max_seq_len, cfg.n_classes, BATCH_SEED, vocab=cfg.vocab_size)
    return torch.from_numpy(x), torch.from_numpy(y)


TRAIN_KW = dict(TINY, class_weight=0.7, recon_weight=1.3)


def train_inputs(cfg, step):
    x, y = synth.token_batch(BATCH, cfg.max_seq_len, cfg.n_classes, BATCH_SEED + 100 + step, vocab=cfg.vocab_size)
    return torch.from_numpy(x), torch.from_numpy(y)


def params(cfg):
    return randomise(O.init_params(cfg, PARAM_SEED, F64), NOISE_SEED)


def main():
    ref_model = tf_shim.load_reference_model()
    data = {}
    for name, kw in VARIANTS:
        cfg = O.OracleConfig(**kw)
        p = params(cfg)
        x, y = inputs(cfg)
        m = build_reference(ref_model, cfg, p, x)
        tar_inp, tar_real = x[:, :-1], x[:, 1:]
        out = m.call(x, tar_inp, False, *m_masks(ref_model, x, tar_inp))
        data[f"{name}/x"], data[f"{name}/y"] = x.numpy(), y.numpy()
        data[f"{name}/embedding"] = out["embedding"].numpy()
        if out.get("class") is not None:
            data[f"{name}/class"] = out["class"].numpy()
            data[f"{name}/class_loss"] = np.float64(m.losses_manager.compute_loss("class", y, out["class"]))
        data[f"{name}/recon"] = out["recon"].numpy()
        data[f"{name}/recon_loss"] = np.float64(m.losses_manager.compute_loss("recon", tar_real, out["recon"]))
        if not cfg.use_continuous_data:
            pr = m.predict(x)
            data[f"{name}/greedy"] = pr["recon"].numpy().astype(np.int64)
            if "class" in pr and pr["class"] is not None:
                data[f"{name}/pred_class"] = pr["class"].numpy().astype(np.int64)
    # three steps of the reference's trainer (models/sketchformer.py:313-359) with dropout 0 and unequal loss weights
    cfg = O.OracleConfig(**TRAIN_KW)
    p = params(cfg)
    x, y = inputs(cfg)
    m = build_reference(ref_model, cfg, {k: v.clone() for k, v in p.items()}, x, dropout_rate=0.0)
    trainable_variables(m)
    for step in range(3):
        xb, yb = train_inputs(cfg, step)
        qm = m.train_on_batch((xb, yb))
        for k, v in qm.items():
            data[f"train/step{step}/{k}"] = np.float64(v)
    out = m.call(x, x[:, :-1], False, *m_masks(ref_model, x, x[:, :-1]))
    data["train/recon_after"] = out["recon"].detach().numpy()
    np.savez_compressed(OUT, **data)
    print(OUT, os.path.getsize(OUT), "bytes;", len(data), "arrays")


if __name__ == "__main__":
    main()

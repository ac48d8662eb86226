"""The oracle against the REFERENCE'S OWN model code: models/sketchformer.py (build_model, call, encode, decode,
classify_from_embedding, encode_from_seq, predict_class, predict / predict_from_embedding incl. make_dummy_input) and
builders/losses.py, executed from /root/reference on oracle/tf_shim.py (see tests/test_reference_layers.py for what
the shim does and does not pin).  build_model() is the reference's: which layers exist under which hparams, the
schedule and the optimizer hyper-parameters come from its code, the weights from the oracle's init_params."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import sketchformer_oracle as O
from oracle import tf_shim
from oracle.make_golden import TINY, TINY_LOWERDIM0, TINY_NOTBLIND, TINY_V2
from sketchformer_b200 import synth
from oracle.reference_model import build_reference, m_masks, randomise as _randomise

pytestmark = pytest.mark.skipif(not tf_shim.available(), reason="/root/reference is not present on this box")
F64 = torch.float64


@pytest.fixture(scope="module")
def ref_model():
    return tf_shim.load_reference_model()


def _batch(cfg, B=4, seed=21):
    if cfg.use_continuous_data:
        x, y = synth.continuous_batch(B, cfg.max_seq_len, cfg.n_classes, seed)
        return torch.from_numpy(x).to(F64), torch.from_numpy(y)
    x, y = synth.token_batch(B, cfg.max_seq_len, cfg.n_classes, seed, vocab=cfg.vocab_size)
    return torch.from_numpy(x), torch.from_numpy(y)


VARIANTS = [("default", TINY), ("v2_classbuffer", dict(TINY_V2)), ("lowerdim0", TINY_LOWERDIM0),
            ("not_blind", TINY_NOTBLIND), ("continuous", dict(TINY, use_continuous_data=True)),
            ("no_recon", dict(TINY, do_reconstruction=False))]


@pytest.mark.parametrize("name,kw", VARIANTS, ids=[v[0] for v in VARIANTS])
def test_forward_and_losses_match_the_reference_model(ref_model, name, kw):
    cfg = O.OracleConfig(**kw)
    p = _randomise(O.init_params(cfg, 13, F64), 14)
    x, y = _batch(cfg)
    m = build_reference(ref_model, cfg, p, x)
    tar_inp, tar_real = x[:, :-1], x[:, 1:]
    got = m.call(x, tar_inp, False, *m_masks(ref_model, x, tar_inp))
    want = O.call(p, x, tar_inp, cfg, False, O.create_masks(x, tar_inp, F64), dtype=F64)
    assert set(k for k in got if got[k] is not None) == set(k for k in ("embedding", "class", "recon") if want.get(k) is not None)
    assert torch.allclose(got["embedding"], want["embedding"], atol=1e-12)
    if cfg.do_classification and cfg.lowerdim:
        assert torch.allclose(got["class"], want["class"], atol=1e-12)          # softmax probabilities
    if cfg.do_reconstruction:
        assert torch.allclose(got["recon"], want["recon"], atol=1e-11), float((got["recon"] - want["recon"]).abs().max())
    # losses: the reference's closures (builders/losses.py) with its weights, against compute_losses
    lo = O.compute_losses(want, tar_real, y, cfg)
    if cfg.do_reconstruction:
        r = m.losses_manager.compute_loss("recon", tar_real, got["recon"])
        assert float(r) == pytest.approx(float(lo["recon_loss"]), rel=1e-12)
    if cfg.do_classification and cfg.lowerdim:
        c = m.losses_manager.compute_loss("class", y, got["class"])
        assert float(c) == pytest.approx(float(lo["class_loss"]), rel=1e-9)      # Keras clips p to [1e-7, 1 - 1e-7]
    assert m.metrics_manager.metric_names == [k for k in ("recon_loss", "recon_acc", "class_loss", "class_acc", "total_loss")
                                              if k in lo or k == "total_loss"]
    # schedule and optimizer as build_model() created them (models/sketchformer.py:111-123)
    assert type(m.learning_rate).__name__ == "WarmupDecay" and m.learning_rate.warmup_steps == 5000
    assert m.optimizer.kwargs == dict(beta_1=0.9, beta_2=0.98, epsilon=1e-9)


@pytest.mark.parametrize("name,kw", [v for v in VARIANTS if v[0] in ("default", "v2_classbuffer", "not_blind")],
                         ids=["default", "v2_classbuffer", "not_blind"])
def test_inference_entry_points_match_the_reference_model(ref_model, name, kw):
    """encode_from_seq / predict_class / predict (greedy decode with make_dummy_input, both cross-mask modes)."""
    cfg = O.OracleConfig(**kw)
    p = _randomise(O.init_params(cfg, 15, F64), 16)
    x, _ = _batch(cfg, B=3, seed=5)
    m = build_reference(ref_model, cfg, p, x)
    sos, eos = cfg.vocab_size - 2, cfg.vocab_size - 1
    got = m.predict_class(x.numpy())
    want = O.predict_class(p, x, cfg, F64)
    assert torch.allclose(got["embedding"], want["embedding"], atol=1e-12)
    assert got["class"].tolist() == want["class"].tolist()
    got = m.predict(x)                      # the reference rounds the embedding to float32 before decoding (line 265)
    want = O.predict(p, x, cfg, sos=sos, eos=eos, dtype=F64)
    assert got["recon"].shape == want["recon"].shape and got["recon"][:, 0].tolist() == [sos] * 3
    assert got["recon"].tolist() == want["recon"].tolist()
    assert got["class"].tolist() == want["class"].tolist()
    assert set(got["attn_weights"]) == set(want["attn_weights"])


def test_continuous_greedy_decode_matches_the_reference_model(ref_model):
    cfg = O.OracleConfig(**dict(TINY, use_continuous_data=True, max_seq_len=12))
    p = _randomise(O.init_params(cfg, 17, F64), 18)
    x, _ = _batch(cfg, B=3, seed=6)
    m = build_reference(ref_model, cfg, p, x)
    emb = O.encode_from_seq(p, x, cfg, F64)["embedding"]
    got = m.predict_from_embedding(emb)
    want = O.predict_from_embedding(p, emb, cfg, dtype=F64)
    assert got["recon"].shape == want["recon"].shape
    assert got["recon"][:, 0].tolist() == [[0., 0., 1., 0., 0.]] * 3
    assert torch.allclose(got["recon"].to(F64), want["recon"], atol=1e-5)       # float32-rounded embedding on one side


def test_train_steps_match_the_reference_trainer(ref_model):
    """model_trainer / train_on_batch (models/sketchformer.py:313-359) executed from the reference for three steps
    with dropout 0: `inp = tar = data`, the slicing of the target, both weighted losses and their sum under the
    gradient tape, apply_gradients over every variable, the running-mean quick metrics.  Gradients are torch
    autograd on both sides and the Adam rule is restated from TensorFlow on both sides (tf_shim.py says so); what is
    pinned is the trainer's wiring and the metric definitions (accuracy over ALL positions, pads included)."""
    from oracle.reference_model import trainable_variables
    cfg = O.OracleConfig(**dict(TINY, class_weight=0.7, recon_weight=1.3))
    p = _randomise(O.init_params(cfg, 19, F64), 20)
    x, y = _batch(cfg, B=4, seed=8)
    m = build_reference(ref_model, cfg, {k: v.clone() for k, v in p.items()}, x, dropout_rate=0.0)
    trainable_variables(m)
    assert sum(v.numel() for v in m.trainable_variables) == O.count_params(p)
    opt = O.AdamTF(p, cfg.d_model)
    run = O.RunningMetrics()
    for step in range(3):
        xb, yb = _batch(cfg, B=4, seed=8 + step)
        got = m.train_on_batch((xb, yb))
        losses, _ = O.train_step(p, opt, xb, yb, cfg, dtype=F64)
        run.update(losses)
        want = run.result()
        assert set(got) == set(want) == {"recon_loss", "recon_acc", "class_loss", "class_acc", "total_loss"}
        for k in want:
            assert float(got[k]) == pytest.approx(float(want[k]), rel=1e-6, abs=1e-9), (step, k)
    # after three optimizer steps (lr(0) = 0, so two effective ones) the weights agree
    out_ref = m.call(x, x[:, :-1], False, *m_masks(ref_model, x, x[:, :-1]))
    out_orc = O.call(p, x, x[:, :-1], cfg, False, O.create_masks(x, x[:, :-1], F64), dtype=F64)
    assert torch.allclose(out_ref["recon"], out_orc["recon"], atol=1e-9)
    assert torch.allclose(out_ref["class"], out_orc["class"], atol=1e-10)
    assert m.optimizer.iterations == 3

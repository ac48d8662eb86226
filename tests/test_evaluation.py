"""Prediction sweeps behind the slow metrics and extract-embeddings (sketchformer_b200/evaluation.py):
host-side layout with a stand-in model here, the real device model in the gpu-marked case."""
import os

import numpy as np
import pytest

from sketchformer_b200 import evaluation as E
from sketchformer_b200.data import get_dataloader_by_name
from sketchformer_b200.tokenizer import Tokenizer
from tests.test_loader import _centroids, _write_dataset


def test_stroke5_conversions_follow_the_reference():
    # utils/sketch.py:161-166, 195-207: argmax one-hot, cut at the first end row, row 0 = end keeps everything
    p = np.array([[0.1, 0.2, 0.7, 0.2, 0.1], [0.3, 0.1, 0.1, 0.8, 0.1], [0.0, 0.0, 0.2, 0.2, 0.6], [9, 9, 1, 0, 0]])
    b = E.to_binary_stroke5(p)
    assert b[:, 2:].tolist() == [[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 0, 0]]
    s = E.to_normal_strokes(b)
    assert s.shape == (2, 3) and s[:, 2].tolist() == [0, 1] and np.allclose(s[:, :2], p[:2, :2].astype(np.float32))
    first_is_end = np.array([[0, 0, 0, 0, 1.0], [1, 1, 1, 0, 0]])
    assert E.to_normal_strokes(first_is_end).shape == (2, 3)
    out = E.predictions_to_sketches([p, p[:2]])
    assert out.dtype == object and out[0].shape == (2, 3) and out[1].shape == (2, 3)


class _FakeModel:
    """predict_class = label parity of the first token; predict = echo (ragged widths per batch)"""
    def __init__(self, ds, bs):
        self.dataset, self.hps, self.calls = ds, {"batch_size": bs}, []

    def predict(self, x):
        self.calls.append(("predict", len(x)))
        return {"recon": np.asarray(x)[:, : 5 + len(self.calls)].astype(np.int32)}

    def predict_class(self, x):
        self.calls.append(("class", len(x)))
        x = np.asarray(x)
        return {"class": (x[:, 1] % 7).astype(np.int32), "embedding": x[:, :4].astype(np.float32)}


def test_sweeps_and_npz_schema(tmp_path):
    d = tmp_path / "data"; d.mkdir()
    _write_dataset(str(d))
    Loader = get_dataloader_by_name("stroke3-distributed")
    ds = Loader(Loader.parse_hparams("max_seq_len=32"), str(d), tokenizer=Tokenizer(centroids=_centroids()))
    m = _FakeModel(ds, 10)
    out = E.compute_predictions_on_validation_set(m)
    x, y, pred_x, pred_y, pred_z, tok, plot_fp, tmp_fp, cont = out
    assert len(out) == 9 and tok is ds.tokenizer and cont is False
    assert x.shape == (25, 32) and pred_x.shape[0] == 25           # 32 requested, the split holds 25
    assert y.shape[0] == 25 and pred_y.shape == (25,) and pred_z.shape == (25, 4)
    assert [c for c in m.calls if c[0] == "class"] == [("class", 10), ("class", 10), ("class", 5)]
    all_x, _ = ds.get_all_data_from("valid")
    assert np.array_equal(pred_y, all_x[:, 1] % 7)
    acc = E.classification_accuracy(out)
    assert acc == float(np.mean(np.asarray(y).reshape(-1) == pred_y))
    assert "{}" in plot_fp and tmp_fp.endswith("converted_{}.png")
    f = E.extract_embeddings(m, str(tmp_path / "emb.npz"), set_type="test", batch_size=16, n_samples_to_reconstruct=6)
    z = np.load(f, allow_pickle=True)
    assert sorted(z.files) == ["embeddings", "pred_y", "recon_sketches", "sketches", "y"]
    assert z["embeddings"].shape == (25, 4) and z["pred_y"].shape == (25,) and z["y"].shape[0] == 25
    assert z["sketches"].shape == (6,) and z["recon_sketches"].shape == (6,) and z["sketches"][0].shape[1] == 3


@pytest.mark.gpu
def test_sweeps_through_the_device_model(tmp_path):
    from sketchformer_b200.model import get_model_by_name
    d = tmp_path / "data"; d.mkdir()
    _write_dataset(str(d))
    Loader, Model = get_dataloader_by_name("stroke3-distributed"), get_model_by_name("sketch-transformer-tf2")
    ds = Loader(Loader.parse_hparams("max_seq_len=32"), str(d), tokenizer=Tokenizer(centroids=_centroids(48)))
    hps = Model.parse_hparams(base="batch_size=8", specific="num_layers=1,d_model=64,dff=128,num_heads=4,lowerdim=32,"
                                                           "dropout_rate=0.0")
    model = Model(hps, ds, str(tmp_path / "out"), "e")
    out = E.compute_predictions_on_test_set(model)
    x, y, pred_x, pred_y, pred_z = out[:5]
    assert pred_y.shape == (25,) and pred_z.shape == (25, 64) and pred_x.shape[0] == 25
    assert pred_x[:, 0].tolist() == [ds.tokenizer.SOS] * 25
    # batching does not change results: one batch of 25 vs the sweep's batches of 8
    allx, _ = ds.get_all_data_from("test")
    whole = model.predict_class(allx)
    assert np.array_equal(whole["class"], pred_y) and np.allclose(whole["embedding"], pred_z, rtol=1e-4, atol=1e-5)
    assert 0.0 <= E.classification_accuracy(out) <= 1.0
    f = E.extract_embeddings(model, str(tmp_path / "emb.npz"), batch_size=16, n_samples_to_reconstruct=4)
    z = np.load(f, allow_pickle=True)
    assert z["embeddings"].shape == (25, 64) and z["recon_sketches"].shape == (4,)

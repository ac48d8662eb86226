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


# ---- against the reference's own code (skipped where /root/reference is absent) ---------------------------------
import importlib.util   # noqa: E402
import sys               # noqa: E402
import types             # noqa: E402

REF = "/root/reference"
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "models", "evaluation_mixin.py")),
                               reason="/root/reference is not present on this box")


def _reference_mixin_and_sketch_utils():
    """models/evaluation_mixin.py and utils/sketch.py executed as they lie in the reference, with inert stand-ins for
    the plotting / TensorFlow imports of utils/sketch.py and for the `utils` package __init__."""
    names = ["tensorflow", "svgwrite", "svglib", "svglib.svglib", "reportlab", "reportlab.graphics", "utils",
             "utils.sketch", "utils.helpers", "ref_evaluation_mixin"]
    saved = {k: sys.modules.get(k) for k in names}
    try:
        for n in ("tensorflow", "svgwrite", "svglib", "reportlab"):
            sys.modules[n] = types.ModuleType(n)
        m = types.ModuleType("svglib.svglib"); m.svg2rlg = None; sys.modules["svglib.svglib"] = m
        m = types.ModuleType("reportlab.graphics"); m.renderPM = None; sys.modules["reportlab.graphics"] = m
        utils = types.ModuleType("utils"); utils.__path__ = [os.path.join(REF, "utils")]
        sys.modules["utils"] = utils
        spec = importlib.util.spec_from_file_location("utils.sketch", os.path.join(REF, "utils", "sketch.py"))
        sk = importlib.util.module_from_spec(spec); sys.modules["utils.sketch"] = sk; spec.loader.exec_module(sk)
        utils.sketch = sk
        utils.helpers = types.SimpleNamespace(get_time_id_str=lambda: "TIME")
        spec = importlib.util.spec_from_file_location("ref_evaluation_mixin", os.path.join(REF, "models", "evaluation_mixin.py"))
        mx = importlib.util.module_from_spec(spec); spec.loader.exec_module(mx)
        return mx.TransformerMetricsMixin, sk
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


class _RefFakeModel:
    """Deterministic stand-in for the device model: what matters here is the batching and the return layout."""

    def __init__(self, ds, batch_size, continuous):
        self.dataset, self.hps = ds, {"batch_size": batch_size}
        self.plots_out_dir, self.tmp_out_dir = "/p", "/t"
        self.continuous = continuous

    def predict_class(self, x):
        x = np.asarray(x)
        flat = x.reshape(len(x), -1).astype(np.float64)
        return {"class": (flat.sum(1).astype(np.int64) % 7).astype(np.int32), "embedding": flat[:, :4] * 0.5}

    def predict(self, x):
        x = np.asarray(x)
        if self.continuous:
            rec = np.concatenate([x[:, :1] * 0 + [0., 0., 1., 0., 0.], x], axis=1)
        else:
            rec = np.concatenate([np.full((len(x), 1), self.dataset.tokenizer.SOS), x], axis=1).astype(np.int32)
        return {"recon": rec, **self.predict_class(x)}


@needs_ref
@pytest.mark.parametrize("continuous", [False, pytest.param(True, marks=pytest.mark.skipif(
    tuple(int(v) for v in np.__version__.split(".")[:2]) >= (1, 24),
    reason="utils/sketch.py:211 builds a ragged np.array, which NumPy >= 1.24 refuses (the reference pins 1.17)"))])
def test_prediction_sweep_matches_the_reference_mixin(tmp_path, continuous):
    """compute_predictions_on_set (models/evaluation_mixin.py:16-53) run from the reference on OUR loader and a fake
    model, against sketchformer_b200.evaluation: same samples (seeded 'valid' subset, whole split in file order),
    same batching, same 9-tuple."""
    Mixin, _ = _reference_mixin_and_sketch_utils()
    _write_dataset(str(tmp_path))
    Loader = get_dataloader_by_name("stroke3-distributed")
    spec = "max_seq_len=32" + (",use_continuous_data=True" if continuous else "")
    ds = Loader(Loader.parse_hparams(spec), str(tmp_path), tokenizer=None if continuous else Tokenizer(centroids=_centroids()))

    class RefModel(_RefFakeModel, Mixin):
        pass

    ref_out = RefModel(ds, 64, continuous).compute_predictions_on_set("test")       # one recon batch (n = 25 <= 64)
    got = E.compute_predictions_on_set(_RefFakeModel(ds, 64, continuous), "test")
    assert len(ref_out) == len(got) == 9
    for i in (1, 3, 4):                                    # all_y, pred_y, pred_z
        assert np.array_equal(np.asarray(ref_out[i]), np.asarray(got[i])), i
    for i in (0, 2):                                       # x and pred_x: arrays, or ragged stroke-3 lists
        assert len(ref_out[i]) == len(got[i])
        for a, b in zip(ref_out[i], got[i]):
            assert np.array_equal(np.asarray(a), np.asarray(b)), i
    assert (ref_out[5] is None) == (got[5] is None) == continuous
    assert ref_out[7] == got[7] and ref_out[8] == got[8]
    assert ref_out[6].replace("TIME", "") == got[6][:3] + got[6][3 + 15:]      # '<dir>/<time id>_{}.svg'
    small = E.compute_predictions_on_set(_RefFakeModel(ds, 8, continuous), "test")    # batches of 8: same results
    assert np.array_equal(small[3], got[3]) and np.array_equal(small[4], got[4])


@needs_ref
def test_stroke5_helpers_match_the_reference():
    """to_binary_stroke5 / to_normal_strokes / predictions_to_sketches (utils/sketch.py:161-211)."""
    _, sk = _reference_mixin_and_sketch_utils()
    rng = np.random.default_rng(3)
    preds = rng.normal(size=(6, 20, 5)).astype(np.float32)
    preds[0, :, 4] = -9                  # never ends
    preds[1, 0, 4] = 9                   # "ends" at row 0: kept whole
    preds[2, 7, 4] = 9
    for p in preds:
        assert np.array_equal(sk.to_binary_stroke5(p.copy()), E.to_binary_stroke5(p.copy()))
        b = sk.to_binary_stroke5(p.copy())
        assert np.array_equal(sk.to_normal_strokes(b), E.to_normal_strokes(b))
    # utils/sketch.py:210-211 is np.array([to_normal_strokes(to_binary_stroke5(p)) for p in preds]): a ragged array that
    # NumPy >= 1.24 refuses to build, so the list is compared element by element with our object array
    want = [sk.to_normal_strokes(sk.to_binary_stroke5(p)) for p in preds.copy()]
    got = E.predictions_to_sketches(preds.copy())
    assert len(want) == len(got) and all(np.array_equal(a, b) for a, b in zip(want, got))
    assert len({len(w) for w in want}) > 1

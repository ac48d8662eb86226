"""Callers on the inference side of the hot path: the prediction sweeps that feed the
reference's slow metrics and the `extract-embeddings` experiment.  They only batch host arrays through
Transformer.predict / predict_class (device work) and keep the reference's return layouts; plots, t-SNE and SVG
rendering are control plane and are not built.

  compute_predictions_on_set     models/evaluation_mixin.py:16-53
  classification_accuracy        metrics/classification.py:6-25   ('val-clas-acc', 'test-clas-acc')
  extract_embeddings             experiments/extract_embeddings.py:24-62
"""
import os
import time

import numpy as np


def to_binary_stroke5(s):
    """utils/sketch.py:161-166: pen state -> one-hot of its argmax."""
    s = np.array(s, dtype=np.float32)
    hot = np.argmax(s[:, 2:], axis=-1)
    s[:, 2:] = 0
    s[np.arange(len(hot)), hot + 2] = 1.0
    return s


def to_normal_strokes(big):
    """utils/sketch.py:195-207: stroke-5 -> stroke-3, cut at the first end-of-sketch row (none, or row 0 -> keep all)."""
    ends = np.flatnonzero(np.argmax(big, axis=1) == 4)
    n = int(ends[0]) if len(ends) and ends[0] != 0 else len(big)
    out = np.zeros((n, 3))
    out[:, :2] = big[:n, :2]
    out[:, 2] = big[:n, 3]
    return out


def predictions_to_sketches(preds):
    """utils/sketch.py:210-211 (ragged -> object array)."""
    out = np.empty(len(preds), dtype=object)
    for i, p in enumerate(preds):
        out[i] = to_normal_strokes(to_binary_stroke5(p))
    return out


def _batches(n, bs):
    for i in range(0, n, bs):
        yield i, min(i + bs, n)


def _sweep(model, set_type, batch_size, n_recon):
    ds = model.dataset
    x, _ = ds.get_n_samples_from("valid", n=n_recon, shuffled=True, seeded=True)     # always 'valid', like the reference
    pred_x = [model.predict(x[i:j])["recon"] for i, j in _batches(len(x), batch_size)]
    all_x, all_y = ds.get_all_data_from(set_type)
    pred_y, pred_z = [], []
    for i, j in _batches(len(all_x), batch_size):
        r = model.predict_class(all_x[i:j])
        pred_y.append(r["class"])
        pred_z.append(r["embedding"])
    width = max(p.shape[1] for p in pred_x)            # decode stops early per batch: pad rows to one width
    pred_x = np.concatenate([np.pad(p, [(0, 0), (0, width - p.shape[1])] + [(0, 0)] * (p.ndim - 2)) for p in pred_x], 0)
    return x, all_y, pred_x, np.concatenate(pred_y, 0), np.concatenate(pred_z, 0)


def compute_predictions_on_set(model, set_type):
    """-> (x, y, pred_x, pred_y, pred_z, tokenizer, plot_filepath, tmp_filepath, continuous): the tuple the
    reference's precomputed-prediction metrics unpack."""
    continuous = bool(model.dataset.hps["use_continuous_data"])
    x, all_y, pred_x, pred_y, pred_z = _sweep(model, set_type, model.hps["batch_size"], 32)
    if continuous:
        pred_x, x = predictions_to_sketches(pred_x), predictions_to_sketches(x)
    time_id = time.strftime("%Y%m%d-%H%M%S")
    plot_filepath = os.path.join(getattr(model, "plots_out_dir", "") or "", "{}_{}.svg".format(time_id, "{}"))
    tmp_filepath = os.path.join(getattr(model, "tmp_out_dir", "") or "", "converted_{}.png")
    tokenizer = None if continuous else model.dataset.tokenizer
    return x, all_y, pred_x, pred_y, pred_z, tokenizer, plot_filepath, tmp_filepath, continuous


def compute_predictions_on_validation_set(model):
    return compute_predictions_on_set(model, "valid")


def compute_predictions_on_test_set(model):
    return compute_predictions_on_set(model, "test")


def classification_accuracy(input_data):
    """'val-clas-acc' / 'test-clas-acc': sklearn accuracy_score(y, pred_y) = fraction of exact matches."""
    _, y, _, pred_y = input_data[:4]
    y, pred_y = np.asarray(y).reshape(-1), np.asarray(pred_y).reshape(-1)
    assert len(y) == len(pred_y)
    return float(np.mean(y == pred_y))


def extract_embeddings(model, target_file="embeddings.npz", set_type="valid", batch_size=256, n_samples_to_reconstruct=32):
    """Writes {y, sketches, recon_sketches, pred_y, embeddings} like the 'extract-embeddings' experiment."""
    x, all_y, pred_x, pred_y, pred_z = _sweep(model, set_type, batch_size, n_samples_to_reconstruct)
    if model.dataset.hps["use_continuous_data"]:
        pred_x, x = predictions_to_sketches(pred_x), predictions_to_sketches(x)
    else:
        tok = model.dataset.tokenizer

        def ragged(seqs):
            out = np.empty(len(seqs), dtype=object)
            for i, s in enumerate(seqs):
                out[i] = tok.decode_single(np.squeeze(s))
            return out
        pred_x, x = ragged(pred_x), ragged(x)
    np.savez(target_file, y=all_y, sketches=x, recon_sketches=pred_x, pred_y=pred_y, embeddings=pred_z)
    return target_file

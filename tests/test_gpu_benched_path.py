"""Parity of the path bench.py times — not a scaled-down stand-in for it.

bench.py measures `Transformer.train_on_batch` (the reference's models/sketchformer.py:351-359 entry) at
BASELINE.json configs[1]: batch 128, max_seq_len 200, d_model 256, dff 1024, 4 layers, 8 heads, through the plugin
API with host numpy batches, the whole step replayed as one CUDA graph on a priority stream with the weight
gradients on a side stream.  At that size the GEMMs run 100 cluster tiles in two waves and split-K weight gradients
over 25,600 rows; none of that is exercised by the batch <= 8 cases of test_gpu_train.py.  This test drives exactly
that path and compares it with the fp64 oracle (oracle/sketchformer_oracle.py, a CPU restatement of
models/sketchformer.py:325-349) on the same batches and weights:

  * losses of every step (one eager, three graph replays) to 1e-3 relative,
  * class accuracy counts exactly, token accuracy counts to one count (see the comment at the assertion),
  * every gradient tensor of the first REPLAYED step to 1e-3 relative L2 (at the engine's ReLU pattern; key biases,
    whose true gradient is zero, against 1e-4 of the global norm — see tests/test_gpu_train.py),
  * the optimizer really runs inside the graph: weights move between replays and the oracle's losses keep matching.
"""
import numpy as np
import pytest
import torch

from oracle import sketchformer_oracle as O
from sketchformer_b200.data import get_dataloader_by_name
from sketchformer_b200.model import get_model_by_name
from tests.test_gpu_train import CFG2, engine_relu_pattern

pytestmark = pytest.mark.gpu
B = 128
TOL = 1e-3


def _model(batch=B, **extra):
    c = CFG2
    Model = get_model_by_name("sketch-transformer-tf2")
    Loader = get_dataloader_by_name("stroke3-synthetic")
    data_hps = Loader.parse_hparams(f"max_seq_len={c['max_seq_len']}")
    model_hps = Model.parse_hparams(base=f"batch_size={batch}",
                                    specific=f"num_layers={c['num_layers']},d_model={c['d_model']},dff={c['dff']},"
                                             f"num_heads={c['num_heads']},lowerdim={c['lowerdim']},dropout_rate=0.0")
    dataset = Loader(data_hps, None, n_classes=c["n_classes"], seed=1234)
    return Model(model_hps, dataset, None, "parity", device="cuda:0", seed=0, **extra), dataset


def _shorten(x, lo, hi, seed):
    """The same batch cut to loader-like lengths: n ~ U{lo..hi} tokens per sketch, [SOS] body [EOS] PAD*.  (The
    synthetic generator never goes below T / 4; sketches from the stroke3-distributed loader are mostly far shorter
    than max_seq_len, and the whole-row attention kernels trim their work to the last unpadded key.)"""
    rng = np.random.default_rng(seed)
    x = x.copy()
    for b in range(len(x)):
        n = int(rng.integers(lo, hi + 1))
        if x[b, n - 1] != 0:
            x[b, n - 1] = 1003
        x[b, n:] = 0
    return x


def test_train_on_batch_cfg2_b128_graph_replay_matches_oracle():
    model, dataset = _model()
    eng = model._train_engine
    ocfg = O.OracleConfig(**CFG2)
    p = O.init_params(ocfg, 0, torch.float64)
    eng.load_params(p)
    opt = O.AdamTF(p, ocfg.d_model)
    it = dataset.batch_iterator("train", B)
    batches = [next(it) for _ in range(4)]
    # the step whose gradients are compared, and the last one, on short sequences
    batches[1] = (_shorten(batches[1][0], 23, 140, 1), batches[1][1])
    batches[3] = (_shorten(batches[3][0], 3, 40, 3), batches[3][1])

    for s, (x, y) in enumerate(batches):
        if s == 2:
            # Steps 0 and 1 run at optimizer iterations 0 and 1: lr(0) = 0 (builders/schedulers.py:25-29) leaves the
            # weights bit-identical on both sides, so the gradients of step 1 — the first REPLAYED step — are
            # compared at identical weights.  (At lr ~ 1e-3 Adam's first updates are ~lr * sign(g): every element
            # whose gradient is rounding noise would move the two sides 1e-3 apart before the comparison.)  From
            # here on both sides jump into the warm-up ramp so that the in-graph optimizer moves the weights
            # measurably; the graph reads the iteration counter from device memory, so this needs no re-capture.
            opt.iterations = 4000
            eng.set_step(4000)
        assert isinstance(x, np.ndarray) and x.shape == (B, ocfg.max_seq_len)       # host buffers, as bench's e2e
        got = model.train_on_batch((x, y))
        torch.cuda.synchronize()
        replayed = getattr(eng, "_graph", None) is not None
        assert replayed == (s >= 1), "step 0 is eager, every later one a graph replay"
        last = eng.last_step_metrics()
        xt, yt = torch.from_numpy(x), torch.from_numpy(y)
        if s == 1:
            # first replayed step: gradients at the engine's ReLU activation pattern
            pattern = engine_relu_pattern(eng, ocfg, B)
            own = {}
            with O.relu_pattern(pattern, record=own):
                want, grads = O.train_step(p, opt, xt, yt, ocfg, dtype=torch.float64)
            flips = sum(int((own[k] != pattern[k]).sum()) for k in pattern)
            units = sum(pattern[k].numel() for k in pattern)
            assert flips <= 1e-3 * units, f"ReLU pattern differs on {flips} of {units} hidden units"
            gg = eng.export_grads(grads)      # the arena still holds this step's gradients after the in-graph Adam
            G = float(torch.sqrt(sum((v.double() ** 2).sum() for v in grads.values())))
            bad = []
            for k, ref in grads.items():
                num = float((gg[k].double() - ref.double()).norm())
                den = max(float(ref.double().norm()), 1e-4 * G)
                if not num / den < TOL:
                    bad.append((k, num / den))
            assert not bad, f"gradients of the replayed step off: {sorted(bad, key=lambda t: -t[1])[:8]}"
            worst = max(float((gg[k].double() - ref.double()).norm()) / max(float(ref.double().norm()), 1e-4 * G)
                        for k, ref in grads.items())
            print(f"cfg2 B=128 graph replay: worst gradient rel. L2 error {worst:.2e}, ReLU flips {flips}/{units}")
        else:
            want, _ = O.train_step(p, opt, xt, yt, ocfg, dtype=torch.float64)
        for k in ("recon_loss", "class_loss", "total_loss"):
            assert abs(last[k] - want[k]) <= TOL * abs(want[k]), (s, k, last[k], want[k])
        # integer counts / (B*Td).  25,472 positions of near-uniform random-init logits: a top-2 gap below the fp32
        # rounding of the logits (~1e-4 of their spread) can occur at a position whose target is one of the two;
        # at most one such count is tolerated, anything else must be exact
        assert abs(last["recon_acc"] - want["recon_acc"]) <= 1.0 / (B * (ocfg.max_seq_len - 1)) + 1e-12
        assert last["class_acc"] == pytest.approx(want["class_acc"], abs=1e-12)
        assert set(got) == {"recon_loss", "recon_acc", "class_loss", "class_acc", "total_loss"}
    assert eng.step_count() == 4002 and model.current_step == 4
    # the optimizer ran inside the graph: weights left their initial values by about lr per element
    moved = eng.export_params({"enc.l0.ffn.w1.k": p["enc.l0.ffn.w1.k"]})["enc.l0.ffn.w1.k"]
    init = O.init_params(ocfg, 0, torch.float64)["enc.l0.ffn.w1.k"]
    assert float((moved.double() - init).abs().mean()) > 1e-4
    # and the running means returned to the caller are means over the four steps
    assert got["total_loss"] == pytest.approx(eng.stats()[6] / 4, rel=1e-12)


def test_resume_runs_one_eager_step_before_capturing(tmp_path):
    """restore_checkpoint_if_exists() leaves current_step large; the first step of the fresh process must still run
    eagerly (lazy function attributes / tensor-map encoders), the second is captured (core/models.py:333-344)."""
    model, dataset = _model(batch=8)
    it = dataset.batch_iterator("train", 8)
    for _ in range(3):
        model.train_on_batch(next(it))
    path = model.save_checkpoint(str(tmp_path / "ckpt-2.pt"))
    fresh, _ = _model(batch=8)
    assert fresh.restore_checkpoint_if_exists(path) and fresh.current_step == 3
    eng = fresh._train_engine
    fresh.train_on_batch(next(it))
    assert getattr(eng, "_graph", None) is None and eng.eager_train_steps == 1
    fresh.train_on_batch(next(it))
    assert getattr(eng, "_graph", None) is not None
    ref = model.train_on_batch(next(dataset.batch_iterator("train", 8)))
    assert np.isfinite(ref["total_loss"])


def test_cfg4_continuous_b128_graph_replay_matches_oracle():
    """bench.py --config cfg4 (BASELINE.json configs[3]: use_continuous_data=True, batch 128, T 200): the same plugin path —
    host batches, one eager step, then CUDA-graph replays — against the fp64 oracle: losses of both steps, class
    accuracy exactly, and every gradient of the REPLAYED step (builders/losses.py:43-66 is a product of two batch
    means, so its gradient kernel needs the stats of the whole batch first: that ordering is inside the graph too)."""
    c = CFG2
    Model = get_model_by_name("sketch-transformer-tf2")
    Loader = get_dataloader_by_name("stroke3-synthetic")
    data_hps = Loader.parse_hparams(f"max_seq_len={c['max_seq_len']},use_continuous_data=True")
    model_hps = Model.parse_hparams(base=f"batch_size={B}",
                                    specific=f"num_layers={c['num_layers']},d_model={c['d_model']},dff={c['dff']},"
                                             f"num_heads={c['num_heads']},lowerdim={c['lowerdim']},dropout_rate=0.0")
    dataset = Loader(data_hps, None, n_classes=c["n_classes"], seed=99)
    model = Model(model_hps, dataset, None, "parity4", device="cuda:0", seed=0)
    eng = model._train_engine
    ocfg = O.OracleConfig(**CFG2, use_continuous_data=True)
    p = O.init_params(ocfg, 0, torch.float64)
    eng.load_params(p)
    opt = O.AdamTF(p, ocfg.d_model)
    it = dataset.batch_iterator("train", B)
    for s in range(2):
        x, y = next(it)
        assert isinstance(x, np.ndarray) and x.shape == (B, ocfg.max_seq_len, 5)
        got = model.train_on_batch((x, y))
        torch.cuda.synchronize()
        assert (getattr(eng, "_graph", None) is not None) == (s >= 1)
        last = eng.last_step_metrics()
        xt, yt = torch.from_numpy(x), torch.from_numpy(y)
        if s == 1:          # optimizer iterations 0 and 1: lr(0) = 0, so both sides still hold the initial weights
            pattern = engine_relu_pattern(eng, ocfg, B)
            own = {}
            with O.relu_pattern(pattern, record=own):
                want, grads = O.train_step(p, opt, xt, yt, ocfg, dtype=torch.float64)
            flips = sum(int((own[k] != pattern[k]).sum()) for k in pattern)
            assert flips <= 1e-3 * sum(pattern[k].numel() for k in pattern)
            gg = eng.export_grads(grads)
            G = float(torch.sqrt(sum((v.double() ** 2).sum() for v in grads.values())))
            bad = []
            for k, ref in grads.items():
                err = float((gg[k].double() - ref.double()).norm()) / max(float(ref.double().norm()), 1e-4 * G)
                if not err < TOL:
                    bad.append((k, err))
            assert not bad, f"gradients of the replayed cfg4 step off: {sorted(bad, key=lambda t: -t[1])[:8]}"
        else:
            want, _ = O.train_step(p, opt, xt, yt, ocfg, dtype=torch.float64)
        for k in ("recon_loss", "class_loss", "total_loss"):
            assert abs(last[k] - want[k]) <= TOL * abs(want[k]), (s, k, last[k], want[k])
        assert last["class_acc"] == pytest.approx(want["class_acc"], abs=1e-12)
        assert set(got) == {"recon_loss", "class_loss", "class_acc", "total_loss"}      # no recon_acc with continuous data


def test_predict_class_cfg2_b128_matches_oracle():
    """bench.py --config fwd: `predict_class` at batch 128 through the plugin (encoder + bottleneck + classifier):
    embeddings to 1e-3, class ids exactly wherever the oracle's top-2 probabilities are not a near tie."""
    model, dataset = _model()
    ocfg = O.OracleConfig(**CFG2)
    p = O.init_params(ocfg, 0, torch.float64)
    model._train_engine.load_params(p)
    x, _ = next(dataset.batch_iterator("train", B))
    x = _shorten(x, 5, 200, 7)
    out = model.predict_class(x)
    want = O.predict_class(p, torch.from_numpy(x), ocfg, torch.float64)
    emb = torch.from_numpy(out["embedding"])
    assert emb.shape == (B, ocfg.d_model)
    err = float((emb.double() - want["embedding"]).norm() / want["embedding"].norm())
    assert err < TOL, err
    probs = O.encode_from_seq(p, torch.from_numpy(x), ocfg, torch.float64)["class"]
    top2 = probs.topk(2, dim=1).values
    spread = probs.max(dim=1).values - probs.min(dim=1).values
    clear = ((top2[:, 0] - top2[:, 1]) > 2e-3 * spread).numpy()         # not a near tie at the fp32 tolerance
    same = out["class"] == want["class"].numpy()
    assert clear.sum() > B // 2 and same[clear].all(), np.argwhere(~same & clear).ravel()
    assert out["class"].dtype == np.int32

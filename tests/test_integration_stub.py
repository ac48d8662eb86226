"""The reference's control flow executed around the plugin ("what calls the path").

* integration/models/sketchformer_b200.py — the one file INTEGRATION.md tells a maintainer to add to the reference tree —
  is imported under the reference's own `core/models.py` (BaseModel), `core/metrics.py`, `models/evaluation_mixin.py`,
  `utils/hparams.py`, `utils/helpers.py` and `metrics/classification.py`, all executed from /root/reference, and the
  reference's `train()` loop (core/models.py:163-197), its reports, checkpoint cadence, resume and the
  evaluate-metrics.py sequence (`compute_metrics_from` / `plot_and_send_notification_for` / `clean_up_tmp_dir`,
  evaluate-metrics.py:53-61) drive it.
* The standalone class (`sketchformer_b200.model.Transformer`) offers the same three evaluate-metrics hooks; they are
  run with the reference's metric objects.

There is no GPU where the reference tree is and no reference tree where the GPU is, so the device engine is replaced by
tests/fake_engine.py here (call protocol and parameter layout of the real one); what these tests pin is host logic.
TensorFlow is absent: `tf.keras.Model` / `tf.Variable` are the minimal stand-ins below, Slack and matplotlib are
recorders."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest

from sketchformer_b200.data import get_dataloader_by_name
from sketchformer_b200.tokenizer import Tokenizer
from tests import fake_engine
from tests.test_loader import _centroids, _write_dataset

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUB = os.path.join(ROOT, "integration", "models", "sketchformer_b200.py")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "core", "models.py")),
                                reason="/root/reference is not present on this box")


class _Variable:
    """tf.Variable as core/models.py uses it: assign / assign_add / numpy and integer arithmetic."""
    def __init__(self, v, name=None, trainable=False):
        self.v = int(v)

    def assign_add(self, d):
        self.v += int(d)

    def assign(self, v):
        self.v = int(v)

    def numpy(self):
        return self.v

    def __index__(self):
        return self.v

    def __int__(self):
        return self.v

    def __add__(self, o):
        return self.v + o

    def __floordiv__(self, o):
        return self.v // o

    def __rsub__(self, o):
        return o - self.v

    def __format__(self, spec):
        return format(self.v, spec)


class _Recorder:
    """Notifyier / PlotManager stand-in: remembers what it was asked to do."""
    log = []

    def __init__(self, *a, **k):
        _Recorder.log.append(("init", type(self).__name__, a))

    def __getattr__(self, name):
        def call(*a, **k):
            _Recorder.log.append((name, a))
            if name == "plot":                       # PlotManager.plot(path, ...) leaves a file behind
                open(a[0], "wb").close()
        return call


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture
def reference(monkeypatch):
    """The reference's host-side modules under their own names, for the duration of a test."""
    names = ["tensorflow", "utils", "utils.hparams", "utils.helpers", "utils.plots", "metrics", "metrics.classification",
             "core", "core.models", "core.metrics", "core.notifyier", "models", "models.evaluation_mixin",
             "models.sketchformer_b200"]
    saved = {k: sys.modules.get(k) for k in names}
    _Recorder.log = []
    try:
        tf = types.ModuleType("tensorflow")
        tf.keras = types.SimpleNamespace(Model=type("Model", (), {"__init__": lambda self, *a, **k: None}))
        tf.Variable = _Variable
        tf.config = types.SimpleNamespace(experimental_run_functions_eagerly=lambda flag: None)
        sys.modules["tensorflow"] = tf
        utils = types.ModuleType("utils"); utils.__path__ = []
        sys.modules["utils"] = utils
        utils.hparams = _load("utils.hparams", os.path.join(REF, "utils", "hparams.py"))
        utils.helpers = _load("utils.helpers", os.path.join(REF, "utils", "helpers.py"))
        utils.plots = types.ModuleType("utils.plots"); utils.plots.PlotManager = type("PlotManager", (_Recorder,), {})
        sys.modules["utils.plots"] = utils.plots
        core = types.ModuleType("core"); core.__path__ = []
        sys.modules["core"] = core
        core.metrics = _load("core.metrics", os.path.join(REF, "core", "metrics.py"))
        noti = types.ModuleType("core.notifyier"); noti.Notifyier = type("Notifyier", (_Recorder,), {})
        sys.modules["core.notifyier"] = noti
        # the metrics package: the reference's registry function over the metric classes that import here
        # (classification.py needs scikit-learn only; the t-SNE / reconstruction plots need matplotlib and svg tools)
        metrics = types.ModuleType("metrics"); metrics.__path__ = []
        sys.modules["metrics"] = metrics
        cls_mod = _load("metrics.classification", os.path.join(REF, "metrics", "classification.py"))

        class Placeholder(core.metrics.HistoryMetric):
            """a slow metric that cannot be imported here: counts its inputs instead of drawing them"""
            input_type = "predictions_on_validation_set"
            last_value_repr = "n/a"

            def compute(self, input_data):
                return len(input_data[0])

        by_name = {c.name: c for c in (cls_mod.PrecomputedValidationAccuracy, cls_mod.PrecomputedTestAccuracy)}
        metrics.build_metric_by_name = lambda name, params: by_name.get(name, Placeholder)(params)
        core.models = _load("core.models", os.path.join(REF, "core", "models.py"))
        models = types.ModuleType("models"); models.__path__ = []
        sys.modules["models"] = models
        models.evaluation_mixin = _load("models.evaluation_mixin", os.path.join(REF, "models", "evaluation_mixin.py"))
        fake_engine.install(monkeypatch)
        stub = _load("models.sketchformer_b200", STUB)
        yield types.SimpleNamespace(stub=stub, core=core, metrics=metrics, utils=utils, cls=cls_mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def _dataset(tmp_path, T=32):
    d = tmp_path / "data"
    d.mkdir()
    _write_dataset(str(d), n_chunks=2, per_chunk=20)              # 40 train samples -> 5 batches of 8 per epoch
    Loader = get_dataloader_by_name("stroke3-distributed")
    return Loader(Loader.parse_hparams(f"max_seq_len={T}"), str(d), seed=3, tokenizer=Tokenizer(centroids=_centroids(48)))


HPS = "num_layers=1,d_model=64,dff=64,num_heads=4,lowerdim=16,dropout_rate=0.0"


def test_integration_md_shows_the_stub_that_is_tested():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    body = open(STUB).read()
    start = body.index("import os")
    assert body[start:].strip() in text, "INTEGRATION.md section 2 must quote integration/models/sketchformer_b200.py"


def test_reference_train_loop_reports_and_checkpoints_drive_the_stub(reference, tmp_path, capsys):
    Model = reference.stub.TransformerB200
    ds = _dataset(tmp_path)
    out = tmp_path / "out"; out.mkdir()
    # the registry surface train.py uses: name, default / parsed hparams, config path (train.py:50-66)
    assert Model.name == "sketch-transformer-b200"
    hps = Model.parse_hparams(base="batch_size=8,num_epochs=2,log_every=2,safety_save=0.4,save_every=1.0,slack_config=",
                              specific=HPS)
    from sketchformer_b200.model import Transformer
    assert list(Model.specific_default_hparams().values().items()) == list(Transformer.specific_default_hparams().values().items())
    assert Model.get_config_filepath(str(out), "r1") == os.path.join(str(out), "sketch-transformer-b200-r1", "config.json")

    model = Model(hps, ds, str(out), "r1")
    assert model.batches_per_epoch == 5 and sorted(model.quick_metrics) == sorted(Transformer.quick_metrics)
    assert sorted(os.listdir(model.out_dir)) == ["plots", "tmp", "weights"]
    eng = fake_engine.FakeEngine.instances[0]

    model.train()                                   # core/models.py:163-197, 10 steps
    assert model.current_step.numpy() == 10 and model.epoch == 2
    steps = [c for c in eng.calls if c[0] == "train_step"]
    odd = [c for e in fake_engine.FakeEngine.instances[1:] for c in e.calls if c[0] == "train_step"]
    # the training plan runs one eager step, then graph replays; a batch cut short at the end of a split
    # (core/data.py:146-172) runs eagerly on a plan of its own size over the same arenas
    assert len(steps) + len(odd) == 10 and steps[0][3] is True and all(c[3] is False for c in steps[1:])
    assert all(c[1] < 8 and c[3] is True for c in odd) and eng.step_count() == 10
    # quick metrics: what train_on_batch returned is what BaseModel recorded, step by step (running means)
    hist = {k: m.history for k, m in model.quick_metrics.items()}
    assert all(len(h) == 10 for h in hist.values())
    assert hist["total_loss"][-1] == pytest.approx(eng.metrics()["total_loss"])
    assert hist["total_loss"][0] == pytest.approx(hist["recon_loss"][0] + hist["class_loss"][0])
    # checkpoints at the reference's cadence: safety every int(0.4 * 5) = 2 steps (two newest kept), fixed every 5
    assert sorted(os.listdir(model.wgt_out_dir)) == ["ckpt-7.pt", "ckpt-9.pt", "step4.pt", "step9.pt"]
    # end of each epoch: slow metrics computed (in their threads) on the mixin's prediction sweep, plotted, notified
    for m in model.slow_metrics.values():
        m.thread.join()
    acc = model.slow_metrics["val-clas-acc"]
    assert len(acc.history) == 2 and 0.0 <= acc.last_value <= 1.0
    x_valid, y_valid = ds.get_all_data_from("valid")
    assert acc.last_value == pytest.approx(float(np.mean(np.asarray(y_valid).reshape(-1) == x_valid[:, 1] % ds.n_classes)))
    kinds = [e[0] for e in _Recorder.log]
    assert kinds.count("notify_with_image") == 2 and kinds.count("plot") == 2 and "notify_with_message" in kinds
    logged = capsys.readouterr().out
    assert "Epoch 1 Batch 4/5" in logged and "|total_loss=" in logged

    # resume (train.py:71-72): the newest checkpoint is the one the reference's rule wrote after step 9
    # ((9 + 1) % 2 == 0 and (9 + 1) % 5 == 0; step 10 saves nothing) — a fresh model restores it and train() runs
    # the one step that is left
    again = Model(hps, ds, str(out), "r1")
    again.restore_checkpoint_if_exists("latest")
    assert again.current_step.numpy() == 9
    e2 = fake_engine.FakeEngine.instances[-1]
    assert e2.step_count() == 9
    n_before = len(fake_engine.FakeEngine.instances)
    count = lambda: sum(1 for e in fake_engine.FakeEngine.instances[n_before - 1:] for c in e.calls if c[0] == "train_step")
    again.train()
    assert count() == 1 and again.current_step.numpy() == 10
    again.train()
    assert count() == 1                             # nothing left
    again.hps["num_epochs"] = 3                     # one more epoch: 5 steps from step 10
    again.train()
    assert again.current_step.numpy() == 15 and count() == 6 and e2.step_count() == 15
    assert "ckpt-13.pt" in os.listdir(again.wgt_out_dir) and "step14.pt" in os.listdir(again.wgt_out_dir)


def test_evaluate_metrics_sequence_on_the_stub(reference, tmp_path):
    """evaluate-metrics.py:53-61 with BaseModel's own compute_metrics_from / plot / clean-up on the stub."""
    Model = reference.stub.TransformerB200
    ds = _dataset(tmp_path)
    out = tmp_path / "out"; out.mkdir()
    model = Model(Model.parse_hparams(base="batch_size=8,slack_config=", specific=HPS), ds, str(out), "e1")
    model.restore_checkpoint_if_exists("latest")            # nothing there: "[Checkpoint] Not found", still step 0
    assert model.current_step.numpy() == 0
    chosen = {m: reference.metrics.build_metric_by_name(m, model.hps) for m in ("val-clas-acc", "test-clas-acc")}
    model.compute_metrics_from(chosen)
    for split, name in (("valid", "val-clas-acc"), ("test", "test-clas-acc")):
        x, y = ds.get_all_data_from(split)
        assert chosen[name].last_value == pytest.approx(float(np.mean(np.asarray(y).reshape(-1) == x[:, 1] % ds.n_classes)))
    assert len([f for f in os.listdir(model.tmp_out_dir) if f.startswith("precomputed_")]) == 2
    model.plot_and_send_notification_for(chosen)
    assert os.path.exists(os.path.join(model.plots_out_dir, "evaluation_plots.png"))
    model.clean_up_tmp_dir()
    assert os.listdir(model.tmp_out_dir) == []


def test_standalone_class_offers_the_same_hooks_to_the_reference_metrics(reference, tmp_path):
    """sketchformer_b200.model.Transformer (no BaseModel underneath) with the reference's metric objects."""
    from sketchformer_b200.model import get_model_by_name
    Model = get_model_by_name("sketch-transformer-tf2")
    ds = _dataset(tmp_path)
    model = Model(Model.parse_hparams(base="batch_size=8", specific=HPS), ds, str(tmp_path), "s1", device="cpu")
    chosen = {m: reference.metrics.build_metric_by_name(m, model.hps) for m in ("val-clas-acc", "test-clas-acc")}
    eng_calls = lambda: sum(len([c for c in e.calls if c[0] == "encode"]) for e in fake_engine.FakeEngine.instances)
    model.compute_metrics_from(chosen)
    n = eng_calls()
    x, y = ds.get_all_data_from("valid")
    assert chosen["val-clas-acc"].last_value == pytest.approx(float(np.mean(np.asarray(y).reshape(-1) == x[:, 1] % ds.n_classes)))
    model.compute_metrics_from({"val-clas-acc": chosen["val-clas-acc"]})      # inside the cache window: no new sweep
    assert eng_calls() == n and len(chosen["val-clas-acc"].history) == 2
    assert len(os.listdir(model.tmp_out_dir)) == 2
    report = model.plot_and_send_notification_for(chosen)
    import json
    rep = json.load(open(report))
    assert rep["val-clas-acc"]["plot_type"] == "lines" and len(rep["val-clas-acc"]["data"]) == 2
    model.clean_up_tmp_dir()
    assert os.listdir(model.tmp_out_dir) == []
    with pytest.raises(AttributeError, match="compute_nonsense"):
        model.gather_data_for_metric("nonsense")

"""CPU tests of the host side: the C-ABI library loads and exports every symbol of include/skf.h,
host-only queries work without a GPU, and the product-side initialiser equals the oracle's."""
import ctypes as C
import os
import re

import pytest
import torch

from oracle import sketchformer_oracle as O
from oracle.make_golden import CFG1, TINY
from sketchformer_b200 import _lib
from sketchformer_b200.engine import make_cfg
from sketchformer_b200.init import init_params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "skf.h")).read()
    return sorted(set(re.findall(r"\b(skf_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/skf.h but not exported by libskf.so"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert lib.skf_version() == 1


def test_ctypes_struct_layouts_match_the_library():
    """include/skf.h structs as compiled vs their ctypes mirrors (also enforced by _lib.load())."""
    lib = _lib.load()
    assert lib.skf_abi_sizeof(0) == C.sizeof(_lib.SkfCfg)
    assert lib.skf_abi_sizeof(1) == C.sizeof(_lib.SkfGemmEpilogue)
    assert lib.skf_abi_sizeof(2) == C.sizeof(_lib.SkfParamEntry)
    assert lib.skf_abi_sizeof(99) == 0


def test_param_table_matches_oracle_param_count():
    lib = _lib.load()
    for kw, B, want in ((CFG1, 8, 2272630), (dict(num_layers=4, d_model=256, dff=1024, num_heads=8, lowerdim=256,
                                                   max_seq_len=200, vocab_size=1004, n_classes=345), 128, 8299989)):
        c = make_cfg(B, kw["max_seq_len"], kw["d_model"], kw["dff"], kw["num_heads"], kw["num_layers"],
                     kw["vocab_size"], kw["n_classes"], kw["lowerdim"])
        ent = (_lib.SkfParamEntry * 512)()
        n = lib.skf_param_table(C.byref(c), ent, 512)
        assert n > 0
        total = sum(ent[i].rows * ent[i].cols for i in range(n))
        assert total == want                      # SURVEY §8(d) parameter counts
        arena = lib.skf_param_count(C.byref(c))
        assert want <= arena <= want + 8 * n      # 32-byte alignment padding only (16 bytes in the bf16 images)
        offs = [(ent[i].offset, ent[i].rows * ent[i].cols) for i in range(n)]
        for (o0, s0), (o1, _) in zip(offs, offs[1:]):
            assert o0 % 8 == 0 and o0 + s0 <= o1
        assert lib.skf_workspace_bytes(C.byref(c)) > 0


def test_bad_configs_are_rejected_with_a_message():
    lib = _lib.load()
    c = make_cfg(8, 64, 130, 512, 8, 4, 1004, 10, 256)     # d_model % heads != 0
    assert lib.skf_param_count(C.byref(c)) < 0
    assert b"num_heads" in lib.skf_last_error()
    c = make_cfg(8, 64, 128, 512, 8, 4, 1002, 10, 256)    # vocab not a multiple of 4
    assert lib.skf_workspace_bytes(C.byref(c)) == 0
    assert b"vocab" in lib.skf_last_error()
    c = make_cfg(8, 64, 128, 512, 8, 4, 0, 10, 256, continuous=True)   # stroke-5 input needs no vocabulary
    assert lib.skf_workspace_bytes(C.byref(c)) > 0
    with pytest.raises(_lib.SkfError):
        _lib.check(lib.skf_set_option(b"nope", 1), "set_option")


def test_product_initialiser_equals_oracle():
    for kw in (TINY, CFG1, dict(TINY, use_continuous_data=True), dict(TINY, attn_version=2, class_buffer_layers=2)):
        a = init_params(kw, 0)
        b = O.init_params(O.OracleConfig(**kw), 0)
        assert list(a) == list(b)
        for k in a:
            assert torch.equal(a[k], b[k]), k


def test_engine_refuses_to_run_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from sketchformer_b200.engine import Engine
    with pytest.raises(_lib.SkfError):
        Engine(make_cfg(4, 24, 64, 128, 4, 2, 68, 10, 32))


# ---- checkpoint cadence and model-side defaults against the reference's core/models.py ------------------------------
import importlib.util   # noqa: E402
import os               # noqa: E402
import sys              # noqa: E402
import types            # noqa: E402

import pytest           # noqa: E402

REF_CORE_MODELS = "/root/reference/core/models.py"


def _reference_base_model():
    """core/models.py executed from the reference with stand-ins for what its imports need at class-definition time
    (tf.keras.Model as a plain base class, the utils / metrics / notifier packages): the methods under test are pure
    Python."""
    names = ["tensorflow", "utils", "utils.hparams", "metrics", "core", "core.metrics", "core.notifyier", "ref_core_models"]
    saved = {k: sys.modules.get(k) for k in names}
    try:
        tf = types.ModuleType("tensorflow")
        tf.keras = types.SimpleNamespace(Model=type("Model", (), {}))
        sys.modules["tensorflow"] = tf
        spec = importlib.util.spec_from_file_location("utils.hparams", "/root/reference/utils/hparams.py")
        hp = importlib.util.module_from_spec(spec); spec.loader.exec_module(hp)
        utils = types.ModuleType("utils"); utils.hparams = hp
        sys.modules["utils"], sys.modules["utils.hparams"] = utils, hp
        sys.modules["metrics"] = types.ModuleType("metrics")
        core = types.ModuleType("core"); core.__path__ = []
        sys.modules["core"] = core
        m = types.ModuleType("core.metrics"); m.QuickMetric = object; sys.modules["core.metrics"] = m
        m = types.ModuleType("core.notifyier"); m.Notifyier = object; sys.modules["core.notifyier"] = m
        spec = importlib.util.spec_from_file_location("ref_core_models", REF_CORE_MODELS)
        mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
        return mod.BaseModel
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


@pytest.mark.skipif(not os.path.exists(REF_CORE_MODELS), reason="/root/reference is not present on this box")
def test_base_hparams_paths_and_checkpoint_cadence_match_the_reference(tmp_path):
    """BaseModel.base_default_hparams / get_config_filepath / save_checkpoint_if_its_time (core/models.py:17-30,
    109-112, 346-358) executed from the reference, against sketchformer_b200.model.Transformer."""
    from sketchformer_b200.model import Transformer
    Base = _reference_base_model()
    assert list(Base.base_default_hparams().values().items()) == list(Transformer.base_default_hparams().values().items())

    class RefNamed(Base):
        name = Transformer.name
    assert RefNamed.get_config_filepath("/out", "x1") == Transformer.get_config_filepath("/out", "x1")

    class Step:                      # tf.Variable stand-in: .numpy() and `+ 1` are all the method uses
        def __init__(self, v):
            self.v = v

        def numpy(self):
            return self.v

        def __add__(self, o):
            return self.v + o

    for safety_save, save_every, bpe in ((.5, 1., 10), (1., 1., 7), (.25, 2., 8), (.5, 1., 3)):
        ref_saves, our_saves = [], []
        ref = types.SimpleNamespace(hps=dict(safety_save=safety_save, save_every=save_every), batches_per_epoch=bpe,
                                    wgt_out_dir="W",
                                    ckpt_manager=types.SimpleNamespace(save=lambda: ref_saves.append(("safety", ref.current_step.v)) or "p"),
                                    ckpt=types.SimpleNamespace(save=lambda file_prefix: ref_saves.append(("fixed", file_prefix))))
        ours = object.__new__(Transformer)
        ours.hps, ours.batches_per_epoch, ours.wgt_out_dir = dict(ref.hps), bpe, str(tmp_path)
        ours.save_checkpoint = lambda path: our_saves.append(path)
        for step in range(1, 4 * bpe + 1):
            ref.current_step = Step(step)
            Base.save_checkpoint_if_its_time(ref)
            ours.current_step = step
            ours.save_checkpoint_if_its_time()
        ref_safety = [s for k, s in ref_saves if k == "safety"]
        ref_fixed = [p for k, p in ref_saves if k == "fixed"]
        our_safety = [int(os.path.basename(p)[5:-3]) for p in our_saves if os.path.basename(p).startswith("ckpt-")]
        our_fixed = [os.path.basename(p)[:-3] for p in our_saves if os.path.basename(p).startswith("step")]
        assert our_safety == ref_safety and ref_safety, (safety_save, save_every, bpe)
        assert our_fixed == [os.path.basename(p) for p in ref_fixed] and ref_fixed      # 'step<N>' file stems


def test_rolling_checkpoints_leave_tensorflow_bundles_alone(tmp_path):
    """weights/ may also hold the reference's own `ckpt-N.index` / `ckpt-N.data-*` tensor bundles (CheckpointManager,
    core/models.py:326-331; the supported flow is restore('latest') from one and keep training): the rotation of our
    rolling `ckpt-N.pt` files must neither choke on those names nor delete them."""
    from sketchformer_b200.model import Transformer
    for f in ("ckpt-5.index", "ckpt-5.data-00000-of-00001", "checkpoint", "ckpt-old.pt.bak"):
        (tmp_path / f).write_bytes(b"tf")
    ours = object.__new__(Transformer)
    ours.hps, ours.batches_per_epoch, ours.wgt_out_dir = dict(safety_save=.5, save_every=100.), 4, str(tmp_path)
    ours.save_checkpoint = lambda path: open(path, "wb").close()
    for step in range(1, 13):
        ours.current_step = step
        ours.save_checkpoint_if_its_time()
    left = sorted(os.listdir(tmp_path))
    assert [f for f in left if f.endswith(".pt")] == ["ckpt-11.pt", "ckpt-9.pt"]
    assert {"ckpt-5.index", "ckpt-5.data-00000-of-00001", "checkpoint", "ckpt-old.pt.bak"} <= set(left)

"""Parse rules of the hparam surface (reference utils/hparams.py:232-344, 464-550) and the registry /
default sets the drivers rely on (train.py:50-66, evaluate-metrics.py:30-48).  CPU only."""
import json

import pytest

from sketchformer_b200.data import get_dataloader_by_name
from sketchformer_b200.hparams import HParams, combine_hparams_into_one, load_config, save_config
from sketchformer_b200.model import Transformer, get_model_by_name


def test_types_follow_defaults():
    h = HParams(a=1, b=2.0, c="x", d=True, e=[1, 2, 3])
    h.parse("a=5,b=3,c=hello,d=false,e=[4,5]")
    assert (h.a, h.b, h.c, h.d, h.e) == (5, 3.0, "hello", False, [4, 5])
    assert isinstance(h.b, float)
    h.parse("d=1")
    assert h.d is True
    h.parse("d=True")
    assert h.d is True and h.e == [4, 5]
    with pytest.raises(ValueError):          # the reference parses `e[1]=9` into {1: 9} and set_hparam refuses it
        h.parse("e[1]=9")
    h.parse("b=-.5e-3")
    assert h.b == -0.0005


@pytest.mark.parametrize("bad", ["zzz=1", "a=1.5", "a=x", "d=maybe", "a=1,a=2", "a", "a=[1,2]", "e=3"])
def test_bad_values_raise(bad):
    h = HParams(a=1, b=2.0, c="x", d=True, e=[1, 2, 3])
    with pytest.raises(ValueError):
        h.parse(bad)


def test_reference_default_sets():
    M = get_model_by_name("sketch-transformer-tf2")
    assert M is Transformer
    spec = M.specific_default_hparams().values()
    # models/sketchformer.py:27-52
    assert spec == dict(num_layers=4, d_model=128, dff=512, num_heads=8, dropout_rate=0.1, lowerdim=256,
                        attn_version=1, do_classification=True, class_weight=1.0, class_buffer_layers=0,
                        class_dropout=0.1, do_reconstruction=True, recon_weight=1.0, blind_decoder_mask=True,
                        is_training=True, optimizer="Adam", lr=0.01, lr_scheduler="WarmupDecay", warmup_steps=10000)
    base = M.base_default_hparams().values()
    # core/models.py:20-30
    assert base == dict(batch_size=128, num_epochs=10, save_every=1., safety_save=.5, autograph=True, log_every=100,
                        notify_every=1000, slack_config="token.secret", goal="No description")
    hps = M.parse_hparams(base="batch_size=64", specific="d_model=256,dff=1024,do_classification=False")
    assert hps.batch_size == 64 and hps.d_model == 256 and hps.do_classification is False
    assert M.get_config_filepath("/out", "x1") == "/out/sketch-transformer-tf2-x1/config.json"
    D = get_dataloader_by_name("stroke3-synthetic")
    # dataloaders/distributed_stroke3.py:15-25
    assert D.default_hparams().values() == dict(
        max_seq_len=200, shuffle_stroke=False, token_type="dictionary", use_continuous_data=False,
        use_absolute_strokes=False, tokenizer_dict_file="prep_data/sketch_token/token_dict.pkl",
        tokenizer_resolution=100, augment_stroke_prob=0.1, random_scale_factor=0.1)


def test_config_round_trip(tmp_path):
    M = get_model_by_name("sketch-transformer-tf2")
    D = get_dataloader_by_name("stroke3-synthetic")
    m = M.parse_hparams(None, "num_layers=2")
    d = D.parse_hparams("max_seq_len=64")
    path = tmp_path / "config.json"
    save_config(str(path), combine_hparams_into_one(m, d))
    assert json.load(open(path))["max_seq_len"] == 64
    fresh = combine_hparams_into_one(M.default_hparams(), D.default_hparams())
    load_config(fresh, str(path))
    assert fresh.num_layers == 2 and fresh.max_seq_len == 64


def test_synthetic_loader_contract():
    D = get_dataloader_by_name("stroke3-synthetic")
    ds = D(D.parse_hparams("max_seq_len=32"), None, n_classes=7)
    x, y = next(ds.batch_iterator("train", 5))
    assert x.shape == (5, 32) and x.dtype.name == "int64" and y.shape == (5, 1)
    assert (x[:, 0] == ds.tokenizer.SOS).all() and ds.tokenizer.VOCAB_SIZE == 1004
    # two ranks of a world of 2 see disjoint halves of the same global batch
    a = next(D(D.parse_hparams("max_seq_len=32"), None, rank=0, world_size=2).batch_iterator("train", 4))
    b = next(D(D.parse_hparams("max_seq_len=32"), None, rank=1, world_size=2).batch_iterator("train", 4))
    g = next(D(D.parse_hparams("max_seq_len=32"), None).batch_iterator("train", 8))
    assert (a[0] == g[0][:4]).all() and (b[0] == g[0][4:]).all()


# ---- against the reference's own code (skipped where /root/reference is absent) ---------------------------------
import importlib.util   # noqa: E402
import os               # noqa: E402

REF_HPARAMS = "/root/reference/utils/hparams.py"
needs_ref = pytest.mark.skipif(not os.path.exists(REF_HPARAMS), reason="/root/reference is not present on this box")


def _reference_hparams_module():
    spec = importlib.util.spec_from_file_location("ref_utils_hparams", REF_HPARAMS)   # numpy / six / json / re only
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


MODEL_DEFAULTS = dict(num_layers=4, d_model=128, dff=512, num_heads=8, dropout_rate=0.1, lowerdim=256, attn_version=1,
                      do_classification=True, class_weight=1.0, class_buffer_layers=0, class_dropout=0.1,
                      do_reconstruction=True, recon_weight=1.0, blind_decoder_mask=True, is_training=True,
                      optimizer="Adam", lr=0.01, lr_scheduler="WarmupDecay", warmup_steps=10000, batch_size=128,
                      save_every=1., goal="No description", layers=[1, 2, 3], names=["a"])
PARSE_CASES = ["d_model=256,dff=1024", "dropout_rate=0.0,lr=1e-3", "do_classification=False,blind_decoder_mask=0",
               "optimizer=sgd,goal=hello world", "save_every=2", "num_layers=6,num_heads=4,lowerdim=0", "is_training=true",
               "class_weight=2,recon_weight=.5", "", "warmup_steps=5e3", "d_model=12.5", "unknown=3",
               "do_classification=maybe", "num_layers=", "lr=abc", "d_model=256,d_model=512", "layers=[4,5]",
               "layers[1]=9", "layers=7", "d_model=[1,2]", "lr=-2.5E-4,goal=a=b", " d_model = 64 ", "d_model=0x10",
               # blanks: none before the first name, free around `=` and after a comma, kept at the end of a string
               " d_model=64", "d_model = 64 ", "d_model=64 , lr=2", "d_model=64,  lr = 2 ,goal= z", "goal= spaced out ,d_model=3",
               "goal=", "goal=,d_model=2", "d_model=6 4", "lr= .5 ", "d_model=1\n", "d_model=1,\nlr=2",
               # bools
               "is_training=True ", "is_training= true", "is_training=2", "is_training=-1", "is_training=0.0", "is_training=TRUE",
               "is_training=yes", "is_training= 1 ",
               # numbers
               "d_model=+7", "lr=1_0", "d_model=1_0", "lr=inf", "lr=1e",
               # lists: elements separated by commas and / or blanks, empty elements dropped
               "layers=[]", "layers=[1, 2]", "layers=[1,,2]", "layers=[ 1 ,2 ]", "names=[a, b]", "names=[a b  c]", "names=[]",
               "layers=[1.5]", "d_model=[3]", "layers=3", "names=[x],names=[y]",
               # index assignments are parsed into {index: value} and then refused by set_hparam
               "layers[0]=2", "layers[ 0 ]=1", "layers [0]=1", "layers=[1],layers[0]=2", "layers[0]=2,layers=[1]", "layers[0]=[1]",
               "d_model[0]=3",
               # malformed clauses and names
               "_x=1", "d_model=64,", "d_model=64,,lr=1", "d_model=3;lr=2", "d_model==3", "goal=[x]", "d_model.x=3", "a.b=1",
               "goal=x,goal=y"]


@needs_ref
@pytest.mark.parametrize("text", PARSE_CASES)
def test_parse_agrees_with_the_reference_hparams_class(text):
    """utils/hparams.py executed as it lies in the reference: same values and value types, or the same exception
    type, for every `k=v` string (train.py --hparams / --base-hparams / --data-hparams)."""
    ref = _reference_hparams_module().HParams
    res = []
    for cls in (ref, HParams):
        try:
            h = cls(**MODEL_DEFAULTS)
            h.parse(text)
            v = h.values()
            res.append({k: (type(v[k]).__name__, v[k]) for k in sorted(v)})
        except Exception as e:          # noqa: BLE001 - the exception TYPE is part of the contract
            res.append(type(e).__name__)
    assert res[0] == res[1], text


@needs_ref
def test_default_hparam_sets_come_from_the_reference_code():
    """Transformer.specific_default_hparams (models/sketchformer.py:24-53) and the loader's default_hparams
    (dataloaders/distributed_stroke3.py:13-26) evaluated from the reference's files."""
    from oracle import tf_shim
    from oracle.make_loader_golden import reference_loader
    RefT, _ = tf_shim.load_reference_model()                  # its HParams stand-in is `dict`
    assert RefT.name == "sketch-transformer-tf2" == Transformer.name
    assert dict(RefT.specific_default_hparams()) == Transformer.specific_default_hparams().values()
    assert list(RefT.quick_metrics) == list(Transformer.quick_metrics)
    ld = reference_loader({}, None)
    import sys
    import types
    utils = types.ModuleType("utils"); utils.hparams = types.SimpleNamespace(HParams=dict)
    saved = sys.modules.get("utils")
    type(ld).default_hparams.__func__.__globals__["utils"] = utils
    try:
        ref_defaults = dict(type(ld).default_hparams())
    finally:
        if saved is not None:
            sys.modules["utils"] = saved
    D = get_dataloader_by_name("stroke3-distributed")
    assert type(ld).name == D.name == "stroke3-distributed"
    assert ref_defaults == D.default_hparams().values()


@needs_ref
def test_config_json_is_byte_identical_to_the_reference(tmp_path):
    """save_config / load_config / combine_hparams_into_one (utils/hparams.py:49-84; train.py:76-77 writes
    config.json, evaluate-metrics.py reads it): the file the reference code writes for the merged model + data
    hparams, byte for byte, and each side loads the other's file."""
    ref = _reference_hparams_module()
    M, D = get_model_by_name("sketch-transformer-tf2"), get_dataloader_by_name("stroke3-distributed")
    ours = combine_hparams_into_one(M.parse_hparams("batch_size=64", "d_model=256,dff=1024,lr=0.001"),
                                    D.parse_hparams("max_seq_len=64,use_continuous_data=True"))
    theirs = ref.combine_hparams_into_one(ref.HParams(**M.default_hparams().values()).parse("batch_size=64,d_model=256,dff=1024,lr=0.001"),
                                          ref.HParams(**D.default_hparams().values()).parse("max_seq_len=64,use_continuous_data=True"))
    assert list(ours.values().items()) == list(theirs.values().items())          # same keys, same order, same values
    a, b = tmp_path / "ours.json", tmp_path / "ref.json"
    save_config(str(a), ours)
    ref.save_config(str(b), theirs, verbose=False)
    assert a.read_bytes() == b.read_bytes()
    fresh = combine_hparams_into_one(M.default_hparams(), D.default_hparams())
    load_config(fresh, str(b))
    assert fresh.values() == ours.values()
    fresh_ref = ref.combine_hparams_into_one(ref.HParams(**M.default_hparams().values()), ref.HParams(**D.default_hparams().values()))
    ref.load_config(fresh_ref, str(a), verbose=False)
    assert fresh_ref.values() == theirs.values()

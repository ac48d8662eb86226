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
    h.parse("d=True,e[1]=9")
    assert h.d is True and h.e == [4, 9]
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

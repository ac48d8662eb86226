"""Greedy decode parity (Transformer.predict / predict_from_embedding, models/sketchformer.py:201-311):
the KV-cached CUDA decode (skf_decode_begin / skf_decode_steps through the C ABI) against the oracle, which
re-runs the whole decoder on the growing prefix exactly like the reference.

Token ids are compared bit-exactly.  A greedy argmax is discontinuous, so if the sequences differ the test
looks at the first differing step and accepts it only when the oracle's own logits of the two candidates
are tied to within 1e-3 of the logit range (the north-star's fp32 tolerance); otherwise it fails.
"""
import os

import numpy as np
import pytest
import torch

from oracle import sketchformer_oracle as O
from oracle.make_golden import CFG1, TINY
from sketchformer_b200 import synth
from tests.gpu_util import rel_err
from tests.test_gpu_train import build

pytestmark = pytest.mark.gpu


def _assert_same_tokens(p, ocfg, emb, got, want, expected_len=None):
    got, want = np.asarray(got), np.asarray(want)
    n = min(got.shape[1], want.shape[1])
    diff = np.argwhere(got[:, :n] != want[:, :n])
    if diff.size == 0:
        assert got.shape == want.shape, (got.shape, want.shape)
        return
    j = int(diff[:, 1].min())                      # first step at which any row differs
    prefix = torch.from_numpy(want[:, :j].astype(np.int64))
    nattn = expected_len if expected_len is not None else j
    dummy = O.make_dummy_input(ocfg, expected_len, nattn, want.shape[0], torch.float64)
    _, combined, dec_pad = O.create_masks(dummy, prefix, torch.float64)
    z = O.decode(p, emb.double(), prefix, dec_pad, combined, ocfg, dtype=torch.float64)["recon"][:, -1]
    for b in np.unique(diff[diff[:, 1] == j][:, 0]):
        zb = z[b]
        gap = abs(float(zb[int(got[b, j])] - zb[int(want[b, j])]))
        assert gap <= 1e-3 * float(zb.max() - zb.min()), f"row {b} step {j}: {got[b, j]} vs {want[b, j]}, gap {gap}"


# cfg2 dimensions (d_model 256, 8 heads of 32, dff 1024, max_seq_len 200; 2 layers and batch 2 so that the oracle's
# O(T^2) re-evaluation of the decoder on the growing prefix, models/sketchformer.py:274-301, finishes in seconds)
CFG2_L2 = dict(num_layers=2, d_model=256, dff=1024, num_heads=8, lowerdim=256, max_seq_len=200, vocab_size=1004,
               n_classes=345)


@pytest.mark.parametrize("backend", [1, 0], ids=["simt", "tcgen05"])
@pytest.mark.parametrize("name,cfg_kw,B", [("tiny", TINY, 4), ("cfg1", CFG1, 8), ("cfg2-l2", CFG2_L2, 2)],
                         ids=["tiny", "cfg1", "cfg2-l2"])
def test_greedy_decode_tokens(name, cfg_kw, B, backend):
    ocfg, eng = build(cfg_kw, B, backend)
    p = O.init_params(ocfg, 3, torch.float64)
    eng.load_params(p)
    sos, eos = ocfg.vocab_size - 2, ocfg.vocab_size - 1
    x, _ = synth.token_batch(B, ocfg.max_seq_len, ocfg.n_classes, 99, vocab=ocfg.vocab_size)
    xt = torch.from_numpy(x)
    want = O.predict(p, xt, ocfg, sos=sos, eos=eos, dtype=torch.float64)
    eng.encode(xt.cuda())
    got = eng.decode(None, None, sos, eos, chunk=7).cpu().numpy()
    assert got.dtype == np.int32 and got[:, 0].tolist() == [sos] * B
    _assert_same_tokens(p, ocfg, want["embedding"], got, want["recon"].numpy())
    assert torch.equal(eng.tensor("cls.argmax", dtype=torch.int32).cpu()[:B], want["class"])
    # from a host embedding (predict_from_embedding): same result
    got2 = eng.decode(want["embedding"].float().cuda().contiguous(), None, sos, eos).cpu().numpy()
    _assert_same_tokens(p, ocfg, want["embedding"], got2, want["recon"].numpy())


@pytest.mark.parametrize("b_pad,b_eos,cols", [(1.0, 2.4, 5), (3.0, 2.3, 25), (3.0, 2.35, 25)])
def test_greedy_decode_stops_at_eos_and_masks_generated_pad(b_pad, b_eos, cols):
    """Output-layer biases chosen (by scanning the oracle) so that (a) rows emit EOS at different early steps
    and the loop stops once ALL have (5 columns), (b) PAD tokens are generated mid-sequence and must be
    masked as keys of later steps (create_masks on the prefix), (c) some rows finish and the rest run on."""
    B = 4
    ocfg, eng = build(TINY, B, 1)
    p = O.init_params(ocfg, 5, torch.float64)
    sos, eos = ocfg.vocab_size - 2, ocfg.vocab_size - 1
    p["out.b"][0] = b_pad
    p["out.b"][eos] = b_eos
    eng.load_params(p)
    x, _ = synth.token_batch(B, ocfg.max_seq_len, ocfg.n_classes, 7, vocab=ocfg.vocab_size)
    xt = torch.from_numpy(x)
    want = O.predict(p, xt, ocfg, sos=sos, eos=eos, dtype=torch.float64)
    assert want["recon"].shape[1] == cols
    eng.encode(xt.cuda())
    got = eng.decode(None, None, sos, eos, chunk=3).cpu().numpy()
    _assert_same_tokens(p, ocfg, want["embedding"], got, want["recon"].numpy())


def test_greedy_decode_expected_len_mask():
    """blind_decoder_mask=False: predict() passes the exact input length, cross-attention keys beyond it are
    masked (make_dummy_input, sketchformer.py:230-253)."""
    B = 4
    ocfg, eng = build(TINY, B, 1, blind_decoder_mask=False)
    p = O.init_params(ocfg, 1, torch.float64)
    eng.load_params(p)
    sos, eos = ocfg.vocab_size - 2, ocfg.vocab_size - 1
    x, _ = synth.token_batch(B, ocfg.max_seq_len, ocfg.n_classes, 11, vocab=ocfg.vocab_size)
    xt = torch.from_numpy(x)
    want = O.predict(p, xt, ocfg, sos=sos, eos=eos, dtype=torch.float64)
    tlen = (xt > 0).sum(-1)
    eng.encode(xt.cuda())
    got = eng.decode(None, tlen, sos, eos).cpu().numpy()
    _assert_same_tokens(p, ocfg, want["embedding"], got, want["recon"].numpy(), expected_len=tlen.float())


@pytest.mark.parametrize("backend", [1, 0], ids=["simt", "tcgen05"])
def test_greedy_decode_growing_cross_mask(backend):
    """blind_decoder_mask=False and NO expected_len (predict_from_embedding as experiments/interpolations_for_mturk.py:78
    calls it): at step i the reference attends the first i + 1 cross keys for every position (sketchformer.py:275), so
    the whole prefix is recomputed every step.  Tokens against the oracle's restatement of that loop; the class
    prediction comes with it; the same through the plugin API."""
    B = 4
    ocfg, eng = build(TINY, B, backend, blind_decoder_mask=False)
    p = O.init_params(ocfg, 2, torch.float64)
    eng.load_params(p)
    sos, eos = ocfg.vocab_size - 2, ocfg.vocab_size - 1
    x, _ = synth.token_batch(B, ocfg.max_seq_len, ocfg.n_classes, 13, vocab=ocfg.vocab_size)
    xt = torch.from_numpy(x)
    emb = O.predict(p, xt, ocfg, sos=sos, eos=eos, dtype=torch.float64)["embedding"]
    want = O.predict_from_embedding(p, emb, ocfg, expected_len=None, sos=sos, eos=eos, dtype=torch.float64)
    eng.encode(xt.cuda())
    got = eng.decode(None, None, sos, eos).cpu().numpy()
    _assert_same_tokens(p, ocfg, emb, got, want["recon"].numpy())
    # it differs from the cached decode with the true lengths (otherwise this test would not see the growing mask)
    with_len = eng.decode(None, (xt > 0).sum(-1), sos, eos).cpu().numpy()
    n = min(with_len.shape[1], got.shape[1])
    assert with_len.shape != got.shape or not np.array_equal(with_len[:, :n], got[:, :n])
    # a second call gives the same tokens (no state left over from the prefix recomputation)
    again = eng.decode(None, None, sos, eos).cpu().numpy()
    assert np.array_equal(again, got)


@pytest.mark.parametrize("backend", [1, 0], ids=["simt", "tcgen05"])
def test_greedy_decode_continuous(backend):
    B = 4
    ocfg, eng = build(TINY, B, backend, use_continuous_data=True)
    p = O.init_params(ocfg, 2, torch.float64)
    eng.load_params(p)
    x, _ = synth.continuous_batch(B, ocfg.max_seq_len, ocfg.n_classes, 13)
    xt = torch.from_numpy(x)
    want = O.predict(p, xt, ocfg, dtype=torch.float64)
    eng.encode(xt.cuda())
    got = eng.decode(None, None).cpu()
    assert tuple(got.shape) == tuple(want["recon"].shape)
    assert got[:, 0].tolist() == [[0., 0., 1., 0., 0.]] * B
    assert rel_err(got, want["recon"]) < (2e-4 if backend == 1 else 1e-3)


def test_golden_tiny_greedy_decode(golden_dir):
    """Committed fixture: weights after three oracle train steps, greedy decode of the first two sketches."""
    g = np.load(os.path.join(golden_dir, "tiny_tokens.npz"))
    ocfg, eng = build(TINY, 2, 0)
    like = O.init_params(ocfg, 0, torch.float64)
    p = {k: torch.from_numpy(g[f"p_after/{k}"]).double().reshape(like[k].shape) for k in like}
    eng.load_params(p)
    x = torch.from_numpy(g["x"][:2])
    eng.encode(x.cuda())
    got = eng.decode(None, None, ocfg.vocab_size - 2, ocfg.vocab_size - 1).cpu().numpy()
    want = g["pred/recon"]
    emb = O.encode_from_seq(p, x, ocfg, torch.float64)["embedding"]
    _assert_same_tokens(p, ocfg, emb, got, want)


def test_model_predict_api():
    """The reference-facing plugin: predict / predict_from_embedding return the reference's keys and types."""
    from sketchformer_b200.data import get_dataloader_by_name
    from sketchformer_b200.model import get_model_by_name
    Model, Loader = get_model_by_name("sketch-transformer-tf2"), get_dataloader_by_name("stroke3-synthetic")
    ds = Loader(Loader.parse_hparams("max_seq_len=32"), None, n_classes=10)
    hps = Model.parse_hparams(base="batch_size=4", specific="num_layers=2,d_model=64,dff=128,num_heads=4,lowerdim=32,"
                                                            "dropout_rate=0.0")
    model = Model(hps, ds, None, "t")
    x, y = next(ds.batch_iterator("valid", 4))
    out = model.predict(x)
    assert set(out) >= {"embedding", "class", "recon", "attn_weights"}
    assert out["recon"].dtype == np.int32 and out["recon"].shape[0] == 4 and 2 <= out["recon"].shape[1] <= 33
    assert out["recon"][:, 0].tolist() == [ds.tokenizer.SOS] * 4
    assert out["class"].dtype == np.int32 and out["class"].shape == (4,)
    out2 = model.predict_from_embedding(out["embedding"])
    assert np.array_equal(out2["recon"], out["recon"]) and np.array_equal(out2["class"], out["class"])
    # one sketch (1-D input) like the interpolation experiment does
    one = model.predict(x[0])
    assert one["recon"].shape[0] == 1


def test_device_tokenizer_matches_host_preprocessing():
    """skf_tokenize_strokes (clip, bbox normalise, nearest centre, SEP/SOS/EOS, cut, PAD in one kernel) against
    the host restatement of DistributedStroke3DataLoader.preprocess + Tokenizer.encode: token ids bit-exact,
    incl. empty-ish, single-point, overflowing and huge-offset sketches."""
    from sketchformer_b200.tokenizer import Tokenizer, preprocess_sketch_host
    rng = np.random.default_rng(5)
    tok = Tokenizer(centroids=rng.uniform(-1, 1, size=(1000, 2)).astype(np.float32))
    T = 200
    sketches = []
    for n in [1, 2, 3, 50, 120, 198, 199, 200, 260, 700] + list(rng.integers(5, 300, size=54)):
        s = np.zeros((int(n), 3), np.float32)
        s[:, :2] = rng.integers(-60, 60, size=(int(n), 2))
        s[:, 2] = rng.random(int(n)) < 0.1
        s[-1, 2] = 1
        sketches.append(s)
    sketches[3][5, 0] = 5000.0                     # clipped to +-1000 before the bounding box
    sketches[4][:, :2] *= 0.001                    # bbox < 1: normalised by 1
    got = tok.encode_batch_device(sketches, T).cpu().numpy()
    want = np.stack([preprocess_sketch_host(s, tok, T) for s in sketches])
    assert got.shape == want.shape == (len(sketches), T)
    assert np.array_equal(got, want), np.argwhere(got != want)[:5]


def test_train_loop_checkpoint_resume(tmp_path):
    """Transformer.train() cadence + restore (core/models.py:163-197, 333-358): 6 steps straight == 3 steps,
    checkpoint, fresh model, restore 'latest', 3 more steps (weights, Adam state, step counter, metrics)."""
    from sketchformer_b200.data import get_dataloader_by_name
    from sketchformer_b200.model import get_model_by_name
    Model, Loader = get_model_by_name("sketch-transformer-tf2"), get_dataloader_by_name("stroke3-synthetic")
    spec = "num_layers=1,d_model=64,dff=128,num_heads=4,lowerdim=32,dropout_rate=0.1"

    def make(out):
        ds = Loader(Loader.parse_hparams("max_seq_len=24"), None, n_classes=10, n_samples=4 * 3, seed=3)
        hps = Model.parse_hparams(base="batch_size=4,num_epochs=2,safety_save=1.0,save_every=1.0,log_every=1", specific=spec)
        return Model(hps, ds, str(out), "r", seed=1)

    a = make(tmp_path / "a")
    assert a.batches_per_epoch == 3
    a._train_engine.set_step(3000)              # a learning rate that moves the weights
    a.train()                                   # 2 epochs x 3 batches
    assert a.current_step == 6 and os.path.exists(a.config_filepath.replace("config.json", "weights"))
    b = make(tmp_path / "b")
    b._train_engine.set_step(3000)
    b.train(max_steps=3)
    files = sorted(os.listdir(b.wgt_out_dir))
    assert "step2.pt" in files and "ckpt-2.pt" in files, files
    c = make(tmp_path / "b")                    # same output dir: resumes from its weights/
    # like the reference ((current_step + 1) % cadence == 0, core/models.py:346-353) the checkpoint of an epoch
    # of 3 batches is written after batch 2
    assert c.restore_checkpoint_if_exists("latest") and c.current_step == 2
    c.dataset = b.dataset                       # keep drawing the same batch stream
    it = b.dataset.batch_iterator("train", 4)
    for _ in range(2):
        next(it)                                # the checkpoint has seen 2 batches; c continues with batches 3..6
    c.dataset.batch_iterator = lambda *args, **kw: it
    c.train()
    assert c.current_step == 6
    pa, pc = a._train_engine.params.cpu(), c._train_engine.params.cpu()
    # split-K weight gradients and the embedding scatter accumulate with atomics, so two runs agree to fp32
    # rounding, not bit for bit; Adam turns a sign flip of a noise-level gradient into one +-lr step (lr ~ 1e-3)
    assert float((pa - pc).norm() / pa.norm()) < 1e-3 and float((pa - pc).abs().max()) < 5e-3
    assert a._train_engine.step_count() == c._train_engine.step_count() == 3006
    ma, mc = a._train_engine.metrics(), c._train_engine.metrics()
    assert all(abs(ma[k] - mc[k]) <= 1e-3 * max(abs(ma[k]), 1e-3) for k in ma), (ma, mc)


@pytest.mark.gpu
def test_tf_checkpoint_round_trip_through_the_model(tmp_path):
    """save_tf_checkpoint -> fresh model -> restore_checkpoint_if_exists(prefix / 'latest'): weights, Adam moments
    and both step counters come back under the reference's checkpoint keys (core/models.py:321-344)."""
    from sketchformer_b200 import tf_checkpoint
    from sketchformer_b200.data import get_dataloader_by_name
    from sketchformer_b200.model import get_model_by_name
    Model, Loader = get_model_by_name("sketch-transformer-tf2"), get_dataloader_by_name("stroke3-synthetic")
    spec = "num_layers=2,d_model=64,dff=128,num_heads=4,lowerdim=32,dropout_rate=0.0"

    def make(out, seed):
        ds = Loader(Loader.parse_hparams("max_seq_len=24"), None, n_classes=10, n_samples=8, seed=3)
        hps = Model.parse_hparams(base="batch_size=4,num_epochs=1,log_every=1", specific=spec)
        return Model(hps, ds, str(out), "r", seed=seed)

    a = make(tmp_path / "a", 1)
    a._train_engine.set_step(3000)
    batches = [b for _, b in zip(range(3), a.dataset.batch_iterator("train", 4))]
    for b in batches[:2]:
        a.train_on_batch(b)
    a.current_step = 2
    prefix = a.save_tf_checkpoint(os.path.join(a.wgt_out_dir, "ckpt-2"))
    keys = tf_checkpoint.read_tensor_bundle(prefix)
    assert "transformer/encoder/enc_layers/1/mha/wq/kernel/.ATTRIBUTES/VARIABLE_VALUE" in keys
    assert "transformer/decoder/dec_layers/0/mha2/dense/bias/.OPTIMIZER_SLOT/optimizer/v/.ATTRIBUTES/VARIABLE_VALUE" in keys
    assert int(keys["optimizer/iter/.ATTRIBUTES/VARIABLE_VALUE"]) == 3002
    c = make(tmp_path / "a", 99)                         # different initial weights, same output directory
    assert c.restore_checkpoint_if_exists("latest") and c.current_step == 2
    ea, ec = a._train_engine, c._train_engine
    assert torch.equal(ea.params, ec.params) and torch.equal(ea.adam_m, ec.adam_m) and torch.equal(ea.adam_v, ec.adam_v)
    assert ec.step_count() == 3002
    a.train_on_batch(batches[2]); c.train_on_batch(batches[2])     # (returned metrics are running means: not compared)
    assert rel_err(ec.params, ea.params) < 1e-4           # two runs of one step differ by the order of their atomics
    pa, pc = a.predict_class(batches[2][0]), c.predict_class(batches[2][0])
    assert np.array_equal(pa["class"], pc["class"])

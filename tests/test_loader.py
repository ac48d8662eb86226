"""Host tests of the loader side of the boundary: dictionary tokenizer (utils/tokenizer.py), the
`stroke3-distributed` chunked-npz loader (dataloaders/distributed_stroke3.py + core/data.py) on a tiny
dataset written here in the reference's on-disk format, and rank sharding."""
import os

import numpy as np
import pytest

from sketchformer_b200.data import get_dataloader_by_name
from sketchformer_b200.tokenizer import Tokenizer, load_centroids, preprocess_sketch_host

REF_PKL = "/root/reference/prep_data/sketch_token/token_dict.pkl"


def _centroids(k=50, seed=0):
    return np.random.default_rng(seed).uniform(-1, 1, size=(k, 2)).astype(np.float32)


def _sketch(rng, n):
    s = np.zeros((n, 3), np.float32)
    s[:, :2] = rng.integers(-40, 40, size=(n, 2))
    s[:, 2] = rng.random(n) < 0.15
    s[-1, 2] = 1
    return s


def _write_dataset(d, n_chunks=3, per_chunk=40, seed=0):
    rng = np.random.default_rng(seed)
    def chunk(n):
        x = np.empty(n, dtype=object)
        for i in range(n):
            x[i] = _sketch(rng, int(rng.integers(3, 60))).astype(np.int16)
        return x, rng.integers(0, 7, size=n)
    for c in range(n_chunks):
        x, y = chunk(per_chunk)
        np.savez(os.path.join(d, f"train_{c:03d}.npz"), x=x, y=y)
    for name in ("valid", "test"):
        x, y = chunk(25)
        np.savez(os.path.join(d, f"{name}.npz"), x=x, y=y)
    np.savez(os.path.join(d, "meta.npz"), n_classes=7, n_samples_train=n_chunks * per_chunk, n_samples_valid=25,
             n_samples_test=25, class_names=np.array([f"c{i}" for i in range(7)]), std=1.0, mean=0.0)


def test_tokenizer_known_answers():
    """utils/tokenizer.py:35-64: ids shifted by one, SEP after pen lifts, SOS/EOS, padding / truncation fix-up."""
    c = np.array([[0, 0], [1, 0], [0, 1], [-1, -1]], np.float32)
    t = Tokenizer(centroids=c)
    assert (t.PAD, t.SEP, t.SOS, t.EOS, t.VOCAB_SIZE) == (0, 5, 6, 7, 8)
    s = np.array([[0.9, 0.1, 0], [0.1, 0.8, 1], [-0.7, -0.9, 0], [0.01, 0.02, 1]], np.float32)
    assert t.encode(s).tolist() == [6, 2, 3, 5, 4, 1, 5, 7]
    assert t.encode(s, seq_len=10).tolist() == [6, 2, 3, 5, 4, 1, 5, 7, 0, 0]
    t6 = Tokenizer(centroids=c, max_seq_len=6)
    assert t6.encode(s).tolist() == [6, 2, 3, 5, 5, 7]          # cut, last two forced to SEP, EOS
    back = t.decode(t.encode(s))
    assert back[:, 2].tolist() == [0, 1, 0, 1] and np.allclose(back[:, :2], c[[1, 2, 3, 0]])
    # exact tie between centres 0 and 1 -> lowest index (sklearn argmin)
    assert t.predict(np.array([[0.5, 0.0]], np.float32)).tolist() == [0]
    assert t.decode([0, 0]).shape == (1, 3)


@pytest.mark.skipif(not os.path.exists(REF_PKL), reason="reference dictionary not present on this box")
def test_centroids_read_from_the_reference_pickle_without_unpickling():
    c = load_centroids(REF_PKL)
    assert c.shape == (1000, 2) and c.dtype == np.float32
    raw = open(REF_PKL, "rb").read()
    assert np.array_equal(c, np.frombuffer(raw[381:8381], "<f4").reshape(1000, 2))
    t = Tokenizer(REF_PKL)
    assert (t.SEP, t.SOS, t.EOS, t.VOCAB_SIZE) == (1001, 1002, 1003, 1004)


GOLDEN_TOK = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tokenizer_ref.npz")


def test_tokenizer_matches_vectors_made_by_the_reference_code():
    """tests/golden/tokenizer_ref.npz was written by oracle/make_tokenizer_golden.py, which EXECUTES the reference's
    utils/tokenizer.py (encode / decode, every max_seq_len regime: variable length, padding, truncation with the
    SEP / EOS fix-up) on seeded float32 sketches with the pickled centroids.  Bit-exact ids, exact decoded rows."""
    g = np.load(GOLDEN_TOK)
    cen, n = g["centroids"], int(g["n"])
    assert cen.shape == (1000, 2)
    for L in g["seq_lens"].tolist():
        t = Tokenizer(centroids=cen, max_seq_len=L)
        for i in range(n):
            want = g[f"enc/{L}/{i}"]
            got = t.encode(g[f"in/{i}"].copy())
            assert got.dtype.kind == "i" and np.array_equal(got, want), (L, i)
            if L:
                assert len(got) == L
            if f"dec/{L}/{i}" in g:
                assert np.array_equal(np.asarray(t.decode(want), np.float64), g[f"dec/{L}/{i}"]), (L, i)
    t0 = Tokenizer(centroids=cen)
    assert np.array_equal(t0.encode(g["in/1"].copy(), seq_len=300), g["enc_pad/0"])
    lens = {len(g[f"enc/0/{i}"]) for i in range(n)}
    assert min(lens) < 16 and max(lens) > 200          # the fixture covers padding and truncation at every length


@pytest.mark.skipif(not os.path.exists(REF_PKL), reason="/root/reference is not present on this box")
def test_tokenizer_golden_is_reproducible_from_the_reference():
    """Re-runs the reference code here and compares with the committed fixture (guards against a stale file)."""
    from oracle.make_tokenizer_golden import SEQ_LENS, reference_tokenizer, sketches
    g = np.load(GOLDEN_TOK)
    sk = sketches()
    assert len(sk) == int(g["n"]) and tuple(g["seq_lens"].tolist()) == SEQ_LENS
    for L in (0, 64):
        t = reference_tokenizer(g["centroids"], L)
        for i in range(0, len(sk), 5):
            assert np.array_equal(sk[i], g[f"in/{i}"])
            assert np.array_equal(np.asarray(t.encode(sk[i].copy())), g[f"enc/{L}/{i}"])


@pytest.mark.skipif(not os.path.exists(REF_PKL), reason="reference dictionary not present on this box")
def test_nearest_centre_agrees_with_sklearn_kmeans_predict():
    """utils/tokenizer.py:38 calls KMeans.predict of the pickled estimator; here a scikit-learn KMeans carrying the
    reference's 1000 centres (lifted from the pickle) against Tokenizer.predict on 40 k points spread like
    normalised sketch deltas.  Disagreements are only tolerated between centres that are tied to 1e-6."""
    from sklearn.cluster import KMeans
    t = Tokenizer(REF_PKL)
    km = KMeans(n_clusters=t.n_clusters)
    km.cluster_centers_ = t.centroids.astype(np.float32)
    km._n_threads = 1
    km.n_features_in_ = 2
    rng = np.random.default_rng(0)
    X = np.concatenate([rng.normal(0, 0.05, size=(20000, 2)), rng.uniform(-1, 1, size=(20000, 2))]).astype(np.float32)
    want, got = km.predict(X), t.predict(X)
    c = t.centroids.astype(np.float64)
    for i in np.nonzero(want != got)[0]:
        da, db = ((X[i] - c[want[i]]) ** 2).sum(), ((X[i] - c[got[i]]) ** 2).sum()
        assert abs(da - db) <= 1e-6 * max(da, db), (i, want[i], got[i], da, db)
    assert (want != got).mean() < 1e-3


GOLDEN_LOADER = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loader_ref.npz")


def test_preprocess_matches_vectors_made_by_the_reference_code(tmp_path):
    """tests/golden/loader_ref.npz was written by oracle/make_loader_golden.py, which EXECUTES the reference's
    DistributedStroke3DataLoader.preprocess (clip, get_bounds normalisation, tokenizer, cut, PAD / stroke-5 rows) on
    seeded raw sketches: clipped gaps, sub-unit bounding boxes, sketches longer than max_seq_len.  Token ids are
    compared bit for bit, stroke-5 rows to float32 rounding (the reference divides in float32 too)."""
    g = np.load(GOLDEN_LOADER)
    n = int(g["n"])
    raw = [g[f"in/{i}"] for i in range(n)]
    _write_dataset(str(tmp_path), n_chunks=1, per_chunk=4)
    Loader = get_dataloader_by_name("stroke3-distributed")
    tok = Tokenizer(centroids=g["centroids"])
    for L in g["seq_lens"].tolist():
        ds = Loader(Loader.parse_hparams(f"max_seq_len={L}"), str(tmp_path), seed=0, tokenizer=tok)
        got = ds.preprocess([r.copy() for r in raw])
        assert got.shape == (n, L) and np.array_equal(got, g[f"tok/{L}"]), L
        for i in range(n):                      # the single-sketch host path is the same function
            assert np.array_equal(preprocess_sketch_host(raw[i].copy(), tok, L), g[f"tok/{L}"][i])
        dc = Loader(Loader.parse_hparams(f"max_seq_len={L},use_continuous_data=True"), str(tmp_path), seed=0)
        gc = dc.preprocess([r.copy() for r in raw], augment=False)
        want = g[f"cont/{L}"]
        assert gc.shape == want.shape == (n, L, 5)
        assert np.array_equal(gc[..., 2:], want[..., 2:])                       # pen / pad flags exactly
        assert np.allclose(gc[..., :2], want[..., :2], rtol=0, atol=1e-7)


@pytest.mark.skipif(not os.path.exists(REF_PKL), reason="/root/reference is not present on this box")
def test_loader_golden_is_reproducible_from_the_reference():
    from oracle.make_loader_golden import BASE_HPS, raw_sketches, reference_loader
    from oracle.make_tokenizer_golden import reference_tokenizer
    g = np.load(GOLDEN_LOADER)
    sk = raw_sketches()
    assert all(np.array_equal(sk[i], g[f"in/{i}"]) for i in range(len(sk)))
    ld = reference_loader(dict(BASE_HPS, max_seq_len=64, use_continuous_data=False), reference_tokenizer(g["centroids"], 0))
    assert np.array_equal(np.asarray(ld.preprocess([s.copy() for s in sk[:10]])), g["tok/64"][:10])


def test_preprocess_matches_reference_steps():
    """clip +-1000, normalise by max(bbox, 1) of the absolute path incl. the origin, tokenise, cut to T, PAD."""
    t = Tokenizer(centroids=_centroids())
    s = np.array([[2000, 0, 0], [-500, 300, 1], [10, 10, 0]], np.float32)
    out = preprocess_sketch_host(s, t, 8)
    # bbox: x path 1000 -> 500 -> 510 (with origin 0): width 1000; y path 0, 300, 310: height 310 -> max_dim 1000
    want = t.encode(np.array([[1.0, 0, 0], [-0.5, 0.3, 1], [0.01, 0.01, 0]], np.float32))
    assert out.tolist() == want.tolist() + [0] * (8 - len(want))
    long = _sketch(np.random.default_rng(1), 30)
    assert len(preprocess_sketch_host(long, t, 16)) == 16 and preprocess_sketch_host(long, t, 16)[0] == t.SOS


def test_distributed_loader_iterates_chunks(tmp_path):
    _write_dataset(str(tmp_path))
    Loader = get_dataloader_by_name("stroke3-distributed")
    hps = Loader.parse_hparams("max_seq_len=32")
    ds = Loader(hps, str(tmp_path), seed=3, tokenizer=Tokenizer(centroids=_centroids()))
    assert ds.n_classes == 7 and ds.n_samples == 120 and len(ds.class_names) == 7
    it = ds.batch_iterator("train", 16, stop_at_end_of_split=False)
    seen, short = 0, 0
    for _ in range(14):                      # > one pass over the 3 chunks: swaps and wrap-around
        x, y = next(it)
        assert x.shape[1:] == (32,) and x.dtype == np.int64 and y.shape == (len(x), 1) and y.dtype == np.int64
        assert (x[:, 0] == ds.tokenizer.SOS).all() and x.max() < ds.tokenizer.VOCAB_SIZE
        short += len(x) < 16                 # like the reference, the batch that hits the end of the split is cut
        seen += len(x)
    assert seen > 120 and short <= 2
    n = sum(len(x) for x, _ in ds.batch_iterator("valid", 10, stop_at_end_of_split=True))
    assert n >= 25                           # the reference's iterator ends on the sample that wraps around
    xa, ya = ds.get_all_data_from("test")
    assert xa.shape == (25, 32) and ya.shape == (25, 1)
    xs, _ = ds.get_n_samples_from("valid", 5, shuffled=True, seeded=14)
    xs2, _ = ds.get_n_samples_from("valid", 5, shuffled=True, seeded=14)
    assert np.array_equal(xs, xs2)


def test_distributed_loader_continuous_rows(tmp_path):
    _write_dataset(str(tmp_path), n_chunks=1)
    Loader = get_dataloader_by_name("stroke3-distributed")
    ds = Loader(Loader.parse_hparams("max_seq_len=24,use_continuous_data=True"), str(tmp_path), seed=1)
    x, y = next(ds.batch_iterator("valid", 8))
    assert x.shape == (8, 24, 5) and x.dtype == np.float32
    assert np.allclose((x[..., 2] + x[..., 3] + x[..., 4])[:, :-1], 1.0)    # one-hot pen state ...
    assert (x[:, -1, 4] == 1).all()                                          # ... the last row is always forced to pad
    assert np.abs(x[..., :2]).max() <= 1.0 + 1e-6


def test_distributed_loader_shards_by_rank(tmp_path):
    _write_dataset(str(tmp_path), n_chunks=4, per_chunk=20)
    Loader = get_dataloader_by_name("stroke3-distributed")
    hps = Loader.parse_hparams("max_seq_len=16")
    tok = Tokenizer(centroids=_centroids())
    a = Loader(hps, str(tmp_path), rank=0, world_size=2, seed=5, tokenizer=tok)
    b = Loader(hps, str(tmp_path), rank=1, world_size=2, seed=5, tokenizer=tok)
    fa, fb = set(a.splits["train"].filepaths), set(b.splits["train"].filepaths)
    assert len(fa) == 2 and len(fb) == 2 and not (fa & fb)            # chunk files dealt round-robin
    _write_dataset(str(tmp_path / "one") if (tmp_path / "one").mkdir() is None else "", n_chunks=1, per_chunk=30)
    a = Loader(hps, str(tmp_path / "one"), rank=0, world_size=2, seed=5, tokenizer=tok)
    b = Loader(hps, str(tmp_path / "one"), rank=1, world_size=2, seed=5, tokenizer=tok)
    a._ready("train"); b._ready("train")
    ia, ib = set(a.splits["train"].current_index_shuffle), set(b.splits["train"].current_index_shuffle)
    assert not (ia & ib) and len(ia | ib) == 30                        # one chunk: residue classes of one order


def test_get_n_samples_picks_the_reference_indices(tmp_path):
    """core/data.py:174-209: file order for shuffled=False (whatever the train shuffle is), and for seeded=True the
    permutation numpy's legacy global generator produces after np.random.seed(1)."""
    _write_dataset(str(tmp_path))
    Loader = get_dataloader_by_name("stroke3-distributed")
    ds = Loader(Loader.parse_hparams("max_seq_len=32"), str(tmp_path), seed=5, tokenizer=Tokenizer(centroids=_centroids()))
    for split in ("valid", "train"):
        allx, ally = ds.get_all_data_from(split)
        n = len(allx)
        cur = ds.splits[split].current
        assert np.array_equal(allx, np.stack(list(cur["x"]))) and np.array_equal(ally[:, 0], cur["y"])   # file order
        x5, y5 = ds.get_n_samples_from(split, 5)
        assert np.array_equal(x5, allx[:5]) and np.array_equal(y5, ally[:5])
        np.random.seed(1)
        want = np.random.permutation(n)[:7]
        np.random.seed()
        xs, ys = ds.get_n_samples_from(split, 7, shuffled=True, seeded=True)
        assert np.array_equal(xs, allx[want]) and np.array_equal(ys, ally[want])
        xs2, _ = ds.get_n_samples_from(split, 7, shuffled=True, seeded=True)
        assert np.array_equal(xs, xs2)


REF_PREP = "/root/reference/prep_data/prepare-distributed-quickdraw.py"


@pytest.mark.skipif(not os.path.exists(REF_PREP), reason="/root/reference is not present on this box")
def test_loader_reads_a_dataset_written_by_the_reference_script(tmp_path, monkeypatch):
    """prep_data/prepare-distributed-quickdraw.py (the writer of the chunked on-disk format: train_XXX.npz, valid.npz,
    test.npz, meta.npz) is RUN from the reference on three synthetic QuickDraw class files; our loader must read its
    output: metadata, every test / valid sample in file order with the right label, every train sample once per
    epoch."""
    import runpy
    rng = np.random.default_rng(4)
    raw = tmp_path / "raw"; raw.mkdir()
    names = ["cat", "dog", "tree"]
    (tmp_path / "classes.txt").write_text("\n".join(names + ["unused"]) + "\n")
    per_class = {}
    for c in names:
        def sketches(n):
            out = np.empty(n, dtype=object)
            for i in range(n):
                out[i] = _sketch(rng, int(rng.integers(3, 40))).astype(np.int16)
            return out
        per_class[c] = {"train": sketches(12), "valid": sketches(4), "test": sketches(6)}
        np.savez(raw / f"{c}.npz", **per_class[c])
    target = tmp_path / "chunked"
    monkeypatch.setattr("sys.argv", ["prepare-distributed-quickdraw.py", "--dataset-dir", str(raw), "--class-list",
                                     str(tmp_path / "classes.txt"), "--n-chunks", "3", "--n-classes", "3",
                                     "--target-dir", str(target)])
    runpy.run_path(REF_PREP, run_name="__main__")
    assert sorted(os.listdir(target)) == ["meta.npz", "test.npz", "train_000.npz", "train_001.npz", "train_002.npz", "valid.npz"]

    tok = Tokenizer(centroids=_centroids())
    Loader = get_dataloader_by_name("stroke3-distributed")
    T = 24
    ds = Loader(Loader.parse_hparams(f"max_seq_len={T}"), str(target), seed=0, tokenizer=tok)
    meta = np.load(target / "meta.npz", allow_pickle=True)
    assert ds.n_classes == 3 and ds.n_samples == 36 == int(meta["n_samples_train"])
    assert list(ds.class_names) == names and ds.scale_factor == float(meta["std"])
    for split, n_per in (("test", 6), ("valid", 4)):
        x, y = ds.get_all_data_from(split)
        want_x = [preprocess_sketch_host(s, tok, T) for c in names for s in per_class[c][split]]
        assert x.shape == (3 * n_per, T) and np.array_equal(x, np.stack(want_x))
        assert y[:, 0].tolist() == [i for i in range(3) for _ in range(n_per)]
    # "one epoch" over the three train chunks as core/data.py:79-106 counts it: 'end_of_split' is raised when the LAST
    # chunk is swapped in (file_cursor reaches n_files at that moment), so the sweep covers the first two chunks and the
    # one sample that completes the running batch - 12 + 12 + 1 - and never a sample twice
    from collections import Counter
    seen = []
    for bx, by in ds.batch_iterator("train", 5, stop_at_end_of_split=True):
        seen += [(tuple(r.tolist()), int(l)) for r, l in zip(bx, by[:, 0])]
    want = Counter((tuple(preprocess_sketch_host(s, tok, T).tolist()), i) for i, c in enumerate(names)
                   for s in per_class[c]["train"])
    assert len(seen) == 25 and not (Counter(seen) - want)
    # an endless iterator does reach every training sketch
    it = ds.batch_iterator("train", 6, stop_at_end_of_split=False)
    pool = Counter()
    for _ in range(18):
        bx, by = next(it)
        pool.update((tuple(r.tolist()), int(l)) for r, l in zip(bx, by[:, 0]))
    assert set(pool) == set(want)


def test_grid_tokenizer_matches_vectors_made_by_the_reference_code():
    """token_type='grid' (utils/tokenizer.py:103-200 + utils/skt_tools.py:55-96, executed by
    oracle/make_tokenizer_golden.py): encode for two resolutions and two max_seq_len regimes, decode of the encoded
    sequences; sketches that end without a pen lift, have none, or sit exactly on the canvas edge."""
    from sketchformer_b200.tokenizer import GridTokenizer
    g = np.load(GOLDEN_TOK)
    n = int(g["grid/n"])
    t100 = GridTokenizer(resolution=100)
    assert (t100.PAD, t100.SEP, t100.SOS, t100.EOS, t100.VOCAB_SIZE) == (0, 10001, 10002, 10003, 10004)
    for res in (100, 16):
        for L in (0, 32):
            t = GridTokenizer(resolution=res, max_seq_len=L)
            for i in range(n):
                s = g[f"grid/in/{i}"]
                want = g[f"grid/enc/{res}/{L}/{i}"]
                assert np.array_equal(t.encode(s.copy()), want), (res, L, i)
                if L == 0:
                    dec = t.decode(want)
                    assert dec.shape == g[f"grid/dec/{res}/{i}"].shape and np.allclose(dec, g[f"grid/dec/{res}/{i}"], atol=1e-12)
    # a decoded sketch re-encodes to the same tokens (cell centres stay in their cells) - for sketches that stay on
    # the canvas: the reference does not clamp, a point outside gets an id outside the vocabulary and decode drops it
    inside = 0
    for i in range(n):
        enc = g[f"grid/enc/100/0/{i}"]
        if enc.min() > 0 and enc.max() <= t100.EOS:
            inside += 1
            assert np.array_equal(t100.encode(t100.decode(enc)), enc)
    assert inside >= n // 2


def test_distributed_loader_with_the_grid_tokenizer(tmp_path):
    """token_type=grid: the loader builds GridTokenizer(resolution=100) (distributed_stroke3.py:35-36) and its batches
    are the per-sketch preprocessing with that tokenizer (vocabulary 10004)."""
    from sketchformer_b200.tokenizer import GridTokenizer
    _write_dataset(str(tmp_path), n_chunks=1, per_chunk=6)
    Loader = get_dataloader_by_name("stroke3-distributed")
    ds = Loader(Loader.parse_hparams("max_seq_len=40,token_type=grid"), str(tmp_path), seed=0)
    assert isinstance(ds.tokenizer, GridTokenizer) and ds.tokenizer.VOCAB_SIZE == 10004
    x, y = ds.get_all_data_from("test")
    raw = np.load(tmp_path / "test.npz", allow_pickle=True)["x"]
    assert x.shape == (25, 40) and x.dtype.kind == "i" and int(x.max()) <= ds.tokenizer.EOS
    for i in (0, 7, 24):
        assert np.array_equal(x[i], preprocess_sketch_host(raw[i], ds.tokenizer, 40))
    assert (x[:, 0] == ds.tokenizer.SOS).all()


@pytest.mark.skipif(not os.path.exists(REF_PKL), reason="/root/reference is not present on this box")
def test_absolute_and_shuffled_strokes_match_the_reference(tmp_path):
    """use_absolute_strokes / shuffle_stroke (distributed_stroke3.py:107-112) through the reference's own preprocess,
    helper by helper and end to end.  np.random.shuffle of the reference is replayed by seeding numpy's global
    generator and shuffling an index list of the same length (same swaps)."""
    from oracle.make_loader_golden import BASE_HPS, raw_sketches, reference_loader
    from oracle.make_tokenizer_golden import reference_tokenizer
    from sketchformer_b200.tokenizer import convert_to_absolute, normalise_sketch, shuffle_strokes
    cen = np.load(GOLDEN_LOADER)["centroids"]
    raw = raw_sketches()[:24]
    ld = reference_loader(dict(BASE_HPS, max_seq_len=64, use_continuous_data=True), None)
    import sys
    sk_mod = type(ld).preprocess.__globals__["utils"].sketch
    tu = type(ld).preprocess.__globals__["utils"].tu_sketch_tools
    for s in raw:
        n = normalise_sketch(s)
        assert np.array_equal(sk_mod.convert_to_absolute(n.copy()), convert_to_absolute(n))
        lines = tu.strokes_to_lines(n.astype(np.float64), scale=1.0, start_from_origin=True)
        perm = np.random.default_rng(len(s)).permutation(len(lines))
        want = tu.lines_to_strokes([lines[i] for i in perm])
        got = shuffle_strokes(n, perm=perm)
        assert got.shape == want.shape and np.allclose(got, want, rtol=0, atol=1e-15)
    # end to end, absolute coordinates: stroke-5 rows and dictionary tokens
    _write_dataset(str(tmp_path), n_chunks=1, per_chunk=4)
    Loader = get_dataloader_by_name("stroke3-distributed")
    tok = Tokenizer(centroids=cen)
    for cont in (True, False):
        hp = dict(BASE_HPS, max_seq_len=64, use_continuous_data=cont, use_absolute_strokes=True)
        ref = reference_loader(hp, None if cont else reference_tokenizer(cen, 0))
        want = np.asarray(ref.preprocess([s.copy() for s in raw], augment=False))
        spec = "max_seq_len=64,use_absolute_strokes=True" + (",use_continuous_data=True" if cont else "")
        ours = Loader(Loader.parse_hparams(spec), str(tmp_path), seed=0, tokenizer=None if cont else tok)
        got = ours.preprocess([s.copy() for s in raw])
        assert got.shape == want.shape
        assert np.allclose(got, want, rtol=0, atol=1e-6) if cont else np.array_equal(got, want)
    # end to end, shuffled strokes (continuous rows): same permutation on both sides, sketch by sketch
    ref = reference_loader(dict(BASE_HPS, max_seq_len=64, use_continuous_data=True, shuffle_stroke=True), None)
    for k, s in enumerate(raw):
        np.random.seed(100 + k)
        want = np.asarray(ref.preprocess([s.copy()], augment=False))[0]
        n_lines = max(1, int((s[:, 2] == 1).sum()))
        np.random.seed(100 + k)
        idx = list(range(n_lines))
        np.random.shuffle(idx)
        np.random.seed()
        n = shuffle_strokes(normalise_sketch(s), perm=idx)[:64]
        assert np.allclose(n[:, :2], want[:len(n), :2], rtol=0, atol=1e-6) and np.array_equal(n[:, 2], want[:len(n), 3])
        assert (want[len(n):, 4] == 1).all()


@pytest.mark.skipif(not os.path.exists(REF_PKL), reason="/root/reference is not present on this box")
def test_point_dropout_augmentation_matches_the_reference():
    """augment_strokes (utils/sketch.py:127-149; train-time augmentation of continuous data): same kept points and
    merged offsets when both sides draw the same uniform numbers (numpy's legacy generator, one draw per point)."""
    from types import SimpleNamespace
    from oracle.make_loader_golden import BASE_HPS, raw_sketches, reference_loader
    from sketchformer_b200.data import augment_strokes
    ld = reference_loader(dict(BASE_HPS, max_seq_len=64, use_continuous_data=True), None)
    ref_aug = type(ld).preprocess.__globals__["utils"].sketch.augment_strokes
    dropped = 0
    for k, s in enumerate(raw_sketches()):
        s = s.astype(np.float32)
        np.random.seed(k)
        want = ref_aug(s.copy(), 0.3)
        np.random.seed()
        got = augment_strokes(s.copy(), 0.3, SimpleNamespace(random=np.random.RandomState(k).rand))
        assert got.shape == want.shape and np.array_equal(got, want.astype(np.float32))
        dropped += len(s) - len(got)
    assert dropped > 50


@pytest.mark.gpu
def test_device_chunk_preprocessing_equals_the_host_loop(tmp_path):
    """load_next_megabatch tokenises whole chunks with skf_tokenize_strokes (one launch per chunk) and the iterator cuts
    batches by fancy indexing: same token ids, bit for bit, and the same batch sequence as the host restatement that is
    pinned to the reference's vectors (tests/golden/loader_ref.npz)."""
    _write_dataset(str(tmp_path), n_chunks=3, per_chunk=70)
    Loader = get_dataloader_by_name("stroke3-distributed")
    hps = Loader.parse_hparams("max_seq_len=48")
    cents = np.random.default_rng(1).uniform(-1, 1, size=(1000, 2)).astype(np.float32)
    dev = Loader(hps, str(tmp_path), seed=11, tokenizer=Tokenizer(centroids=cents))
    host = Loader(hps, str(tmp_path), seed=11, tokenizer=Tokenizer(centroids=cents), device=None)
    assert dev._device_path() and not host._device_path()
    raw = np.load(os.path.join(str(tmp_path), "train_000.npz"), allow_pickle=True)["x"]
    assert np.array_equal(dev.preprocess(raw), host.preprocess(raw))
    a, b = dev.batch_iterator("train", 16), host.batch_iterator("train", 16)
    for _ in range(40):                       # three passes: chunk swaps, the cut batch at the end of the split
        (xa, ya), (xb, yb) = next(a), next(b)
        assert xa.dtype == np.int64 and np.array_equal(xa, xb) and np.array_equal(ya, yb)
    # device-resident batches: the same ids as CUDA tensors
    import torch
    res = Loader(hps, str(tmp_path), seed=11, tokenizer=Tokenizer(centroids=cents), device_batches=True)
    xr, yr = next(res.batch_iterator("train", 16))
    xh, yh = next(Loader(hps, str(tmp_path), seed=11, tokenizer=Tokenizer(centroids=cents), device=None).batch_iterator("train", 16))
    assert torch.is_tensor(xr) and xr.is_cuda and np.array_equal(xr.cpu().numpy(), xh) and np.array_equal(yr, yh)

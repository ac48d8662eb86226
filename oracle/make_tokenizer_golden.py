"""Golden vectors for the dictionary tokenizer made by the REFERENCE's own code (test infrastructure).

    python -m oracle.make_tokenizer_golden          # needs /root/reference; writes tests/golden/tokenizer_ref.npz

The reference class `utils/tokenizer.py:16-101` is executed as it lies under /root/reference.  Its package
`utils/__init__.py` imports plotting libraries that are not installed, so the module is loaded on its own with
empty stand-ins for the two sibling modules it imports but does not use in encode / decode
(`utils.helpers.load_pickle`, `utils.skt_tools`).  The pickled scikit-learn 0.21 estimator cannot be unpickled by
scikit-learn 1.9 (module `sklearn.cluster.k_means_` is gone), so the instance is created without `__init__` and its
`dict` is a scikit-learn KMeans carrying the pickled `cluster_centers_` (lifted from the pickle stream by
`sketchformer_b200.tokenizer.load_centroids`): `dict.predict`, `dict.n_clusters` and `dict.cluster_centers_` are all
the reference code touches.  Inputs are seeded; the fixture stores the centroids, the stroke-3 inputs and the
reference's outputs, so the parity tests run where /root/reference is absent (the GPU box).
"""
import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
PKL = os.path.join(REF, "prep_data/sketch_token/token_dict.pkl")
OUT = os.path.join(ROOT, "tests", "golden", "tokenizer_ref.npz")
SEQ_LENS = (0, 16, 64, 200)


def load_reference_tokenizer_class(which="Tokenizer"):
    """utils/tokenizer.py executed from the reference.  `utils.skt_tools` (strokes_to_lines / lines_to_strokes, used
    by GridTokenizer) is the reference's own file too, loaded with empty stand-ins for the drawing libraries it
    imports at the top (svgwrite, PIL, rdp, svgpathtools)."""
    pkg = types.ModuleType("utils")
    pkg.__path__ = [os.path.join(REF, "utils")]
    helpers = types.ModuleType("utils.helpers")
    helpers.load_pickle = lambda path: (_ for _ in ()).throw(RuntimeError("not used"))
    names = ("utils", "utils.helpers", "utils.skt_tools", "utils.tokenizer", "svgwrite", "PIL", "rdp", "svgpathtools")
    saved = {k: sys.modules.get(k) for k in names}
    sys.modules.update({"utils": pkg, "utils.helpers": helpers})
    try:
        sys.modules["svgwrite"] = types.ModuleType("svgwrite")
        pil = types.ModuleType("PIL"); pil.Image = pil.ImageDraw = None; sys.modules["PIL"] = pil
        rdp = types.ModuleType("rdp"); rdp.rdp = None; sys.modules["rdp"] = rdp
        svp = types.ModuleType("svgpathtools"); svp.real = svp.imag = svp.svg2paths = svp.wsvg = None
        sys.modules["svgpathtools"] = svp
        spec = importlib.util.spec_from_file_location("utils.skt_tools", os.path.join(REF, "utils", "skt_tools.py"))
        skt = importlib.util.module_from_spec(spec)
        sys.modules["utils.skt_tools"] = skt
        spec.loader.exec_module(skt)
        spec = importlib.util.spec_from_file_location("utils.tokenizer", os.path.join(REF, "utils", "tokenizer.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["utils.tokenizer"] = mod
        spec.loader.exec_module(mod)
        return getattr(mod, which)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def reference_tokenizer(centroids, max_seq_len):
    from sklearn.cluster import KMeans
    cls = load_reference_tokenizer_class()
    km = KMeans(n_clusters=len(centroids))
    km.cluster_centers_ = np.asarray(centroids, np.float32)
    km._n_threads = 1
    km.n_features_in_ = 2
    t = object.__new__(cls)
    t.max_seq_len = max_seq_len
    t.dict = km
    t.PAD, t.SEP, t.SOS, t.EOS = 0, km.n_clusters + 1, km.n_clusters + 2, km.n_clusters + 3
    t.VOCAB_SIZE = km.n_clusters + 4
    return t


def sketches(n=48, seed=7):
    """Normalised float32 stroke-3 sketches of ragged length (1 ... 260 points: shorter and longer than every
    SEQ_LEN)."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        m = int(rng.integers(1, 40)) if i % 3 else int(rng.integers(150, 260))
        s = np.zeros((m, 3), np.float32)      # float32 as the loader hands them over (distributed_stroke3.py:94-104)
        s[:, :2] = rng.normal(0, 0.06, size=(m, 2))
        s[:, 2] = rng.random(m) < 0.12
        if i % 5 == 0:
            s[-1, 2] = 1
        out.append(s)
    return out


def grid_sketches(n=30, seed=19):
    """Normalised sketches whose absolute path stays inside the canvas most of the time; some end without a pen lift
    (their tail is dropped by the reference), some have none at all, some touch the upper edge exactly."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        m = int(rng.integers(2, 30))
        s = np.zeros((m, 3), np.float32)
        s[:, :2] = rng.normal(0, 0.12, size=(m, 2))
        s[0, :2] = rng.uniform(-0.5, 0.5, size=2)
        s[:, 2] = rng.random(m) < 0.2
        if i % 3 == 0:
            s[-1, 2] = 1
        if i % 7 == 0:
            s[:, 2] = 0
        if i % 10 == 0:
            s[:, :2] = 0
            s[0, :2] = (1.0, -1.0)          # lands exactly on the upper x edge and the lower y edge
        out.append(s)
    return out


def main():
    from sketchformer_b200.tokenizer import load_centroids
    cen = load_centroids(PKL)
    data = {"centroids": cen, "seq_lens": np.array(SEQ_LENS)}
    sk = sketches()
    data["n"] = np.array(len(sk))
    for i, s in enumerate(sk):
        data[f"in/{i}"] = s
    for L in SEQ_LENS:
        t = reference_tokenizer(cen, L)
        for i, s in enumerate(sk):
            enc = t.encode(s.copy())
            data[f"enc/{L}/{i}"] = np.asarray(enc, np.int64)
            if L in (0, 64):
                data[f"dec/{L}/{i}"] = np.asarray(t.decode(enc), np.float64)
    # seq_len argument (padding of a variable-length tokenizer) and list decoding
    t0 = reference_tokenizer(cen, 0)
    data["enc_pad/0"] = np.asarray(t0.encode(sk[1].copy(), seq_len=300), np.int64)
    # token_type='grid' (utils/tokenizer.py:103-200).  Inputs go in as float64: the reference accumulates the absolute
    # path in Python floats (NumPy 1.17 promotes float32 + int to float64), which NumPy 2 would otherwise do in float32
    Grid = load_reference_tokenizer_class("GridTokenizer")
    gsk = grid_sketches()
    data["grid/n"] = np.array(len(gsk))
    for i, s in enumerate(gsk):
        data[f"grid/in/{i}"] = s
    for res in (100, 16):
        for L in (0, 32):
            t = Grid(resolution=res, max_seq_len=L)
            for i, s in enumerate(gsk):
                enc = np.asarray(t.encode(s.astype(np.float64)), np.int64)
                data[f"grid/enc/{res}/{L}/{i}"] = enc
                if L == 0:
                    data[f"grid/dec/{res}/{i}"] = np.asarray(t.decode(enc), np.float64)
    np.savez_compressed(OUT, **data)
    print(OUT, os.path.getsize(OUT), "bytes;", len(sk), "sketches x", len(SEQ_LENS), "max_seq_len settings")


if __name__ == "__main__":
    main()

"""Golden vectors for the loader's per-sketch preprocessing made by the REFERENCE's own code (test infrastructure).

    python -m oracle.make_loader_golden           # needs /root/reference; writes tests/golden/loader_ref.npz

Executes `DistributedStroke3DataLoader.preprocess` / `_cap_pad_and_convert_sketch`
(dataloaders/distributed_stroke3.py:90-125, 138-153) together with `utils.sketch.get_bounds`
(utils/sketch.py:31-50) and the reference Tokenizer (see oracle/make_tokenizer_golden.py) as they lie under
/root/reference.  The modules are loaded one by one with inert stand-ins for what this image lacks and the
executed code never touches: tensorflow / svgwrite / svglib / reportlab (imported at the top of utils/sketch.py for
its drawing helpers), the `utils` package __init__ (plotting, Slack) and `core.data` (base class with the
prefetch machinery).  The loader instance is created without `__init__` (which globs a dataset directory) and given
exactly the attributes `preprocess` reads: `limit`, `hps`, `tokenizer`.
Raw inputs are seeded integer-valued stroke-3 sketches like the QuickDraw chunks (int16 deltas, pen bit), some with
gaps beyond the +-1000 clip and some longer than max_seq_len.
"""
import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_tokenizer_golden import PKL, REF, reference_tokenizer      # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "loader_ref.npz")
SEQ_LENS = (16, 64, 200)


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def reference_loader(hps, tokenizer):
    """An instance of the reference's DistributedStroke3DataLoader that can run preprocess()."""
    names = ["tensorflow", "svgwrite", "svglib", "svglib.svglib", "reportlab", "reportlab.graphics", "utils",
             "utils.sketch", "utils.skt_tools", "PIL", "rdp", "svgpathtools", "core", "core.data",
             "dataloaders_ref_stroke3"]
    saved = {k: sys.modules.get(k) for k in names}
    try:
        for n in ("tensorflow", "svgwrite", "svglib", "reportlab"):
            sys.modules[n] = types.ModuleType(n)
        m = types.ModuleType("svglib.svglib"); m.svg2rlg = None; sys.modules["svglib.svglib"] = m
        m = types.ModuleType("reportlab.graphics"); m.renderPM = None; sys.modules["reportlab.graphics"] = m
        utils = types.ModuleType("utils")
        utils.__path__ = [os.path.join(REF, "utils")]
        sys.modules["utils"] = utils
        utils.sketch = _load("utils.sketch", os.path.join(REF, "utils", "sketch.py"))
        pil = types.ModuleType("PIL"); pil.Image = pil.ImageDraw = None; sys.modules["PIL"] = pil
        m = types.ModuleType("rdp"); m.rdp = None; sys.modules["rdp"] = m
        m = types.ModuleType("svgpathtools"); m.real = m.imag = m.svg2paths = m.wsvg = None; sys.modules["svgpathtools"] = m
        # utils/__init__.py:5 imports skt_tools under this name; preprocess uses it for shuffle_stroke
        utils.tu_sketch_tools = _load("utils.skt_tools", os.path.join(REF, "utils", "skt_tools.py"))
        core = types.ModuleType("core"); core.__path__ = []
        data = types.ModuleType("core.data")
        data.BaseDataLoader = type("BaseDataLoader", (), {})
        data.DatasetSplit = type("DatasetSplit", (), {})
        sys.modules["core"], sys.modules["core.data"] = core, data
        mod = _load("dataloaders_ref_stroke3", os.path.join(REF, "dataloaders", "distributed_stroke3.py"))
        ld = object.__new__(mod.DistributedStroke3DataLoader)
        ld.limit = 1000                       # distributed_stroke3.py:30
        ld.hps = dict(hps)
        ld.tokenizer = tokenizer
        return ld
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def raw_sketches(n=40, seed=11):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        m = int(rng.integers(2, 50)) if i % 4 else int(rng.integers(120, 240))
        s = np.zeros((m, 3), np.int16)
        s[:, :2] = rng.integers(-60, 60, size=(m, 2))
        if i % 7 == 0:                                   # a gap far beyond the clip limit
            s[int(rng.integers(0, m)), 0] = 2500
            s[int(rng.integers(0, m)), 1] = -1800
        if i % 9 == 0:                                   # a tiny sketch: bounding box below 1 -> divisor 1
            s[:, :2] = 0
            s[0, 0] = 1
        s[:, 2] = rng.random(m) < 0.1
        s[-1, 2] = 1
        out.append(s)
    return out


BASE_HPS = dict(shuffle_stroke=False, use_absolute_strokes=False, token_type="dictionary", augment_stroke_prob=0.1,
                random_scale_factor=0.1)


def main():
    from sketchformer_b200.tokenizer import load_centroids
    cen = load_centroids(PKL)
    sk = raw_sketches()
    data = {"centroids": cen, "seq_lens": np.array(SEQ_LENS), "n": np.array(len(sk))}
    for i, s in enumerate(sk):
        data[f"in/{i}"] = s
    for L in SEQ_LENS:
        ld = reference_loader(dict(BASE_HPS, max_seq_len=L, use_continuous_data=False), reference_tokenizer(cen, 0))
        data[f"tok/{L}"] = np.asarray(ld.preprocess([s.copy() for s in sk], augment=False), np.int64)
        ld = reference_loader(dict(BASE_HPS, max_seq_len=L, use_continuous_data=True), None)
        data[f"cont/{L}"] = np.asarray(ld.preprocess([s.copy() for s in sk], augment=False), np.float64)
    np.savez_compressed(OUT, **data)
    print(OUT, os.path.getsize(OUT), "bytes;", {k: v.shape for k, v in data.items() if k.startswith(("tok", "cont"))})


if __name__ == "__main__":
    main()

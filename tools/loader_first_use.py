"""GPU check: the loader's device tokeniser as the FIRST CUDA work of a fresh process (prefetch threads start inside the
constructor) against the host loop; run it many times: for i in $(seq 20); do python tools/loader_first_use.py; done"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_loader import _write_dataset  # noqa: E402

from sketchformer_b200.data import get_dataloader_by_name  # noqa: E402
from sketchformer_b200.tokenizer import Tokenizer  # noqa: E402

d = tempfile.mkdtemp()
_write_dataset(d, n_chunks=3, per_chunk=70)
cents = np.random.default_rng(1).uniform(-1, 1, size=(1000, 2)).astype(np.float32)
Loader = get_dataloader_by_name("stroke3-distributed")
hps = Loader.parse_hparams("max_seq_len=48")
dev = Loader(hps, d, seed=11, tokenizer=Tokenizer(centroids=cents))
host = Loader(hps, d, seed=11, tokenizer=Tokenizer(centroids=cents), device=None)
raw = np.load(os.path.join(d, "train_000.npz"), allow_pickle=True)["x"]
ok = np.array_equal(dev.preprocess(raw), host.preprocess(raw))
a, b = dev.batch_iterator("train", 16), host.batch_iterator("train", 16)
bad = []
for i in range(14):
    (xa, ya), (xb, yb) = next(a), next(b)
    if not np.array_equal(xa, xb):
        bad.append(i)
print("first-use", "OK" if ok and not bad else f"MISMATCH main={ok} batches={bad}", flush=True)

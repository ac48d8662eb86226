"""GPU diagnostic for the chunk tokeniser (gpurun -- python tools/loader_diag.py): where do device and host differ?"""
import os
import sys
import tempfile
import threading

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_loader import _write_dataset  # noqa: E402

from sketchformer_b200.data import get_dataloader_by_name  # noqa: E402
from sketchformer_b200.tokenizer import Tokenizer, preprocess_sketch_host  # noqa: E402

d = tempfile.mkdtemp()
_write_dataset(d, n_chunks=3, per_chunk=70)
cents = np.random.default_rng(1).uniform(-1, 1, size=(1000, 2)).astype(np.float32)
T = 48
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"


def host_chunk(x):
    tk = Tokenizer(centroids=cents)
    return np.stack([preprocess_sketch_host(s, tk, T) for s in x])


def report(tag, got, want):
    bad = np.flatnonzero((got != want).any(axis=1))
    print(f"{tag}: {len(bad)} of {len(want)} rows differ", bad[:10].tolist(), flush=True)
    return len(bad)


chunks = {f: np.load(os.path.join(d, f), allow_pickle=True)["x"] for f in sorted(os.listdir(d)) if not f.startswith("meta")}
want = {f: host_chunk(x) for f, x in chunks.items()}

# 1. main thread, default stream
tk = Tokenizer(centroids=cents)
for f, x in chunks.items():
    got = tk.encode_batch_device(list(x), T, "cuda:0").cpu().numpy()
    report("main/default " + f, got, want[f])
# 2. main thread, side stream
st = torch.cuda.Stream("cuda:0")
for f, x in chunks.items():
    with torch.cuda.stream(st):
        got = tk.encode_batch_device(list(x), T, "cuda:0").cpu().numpy()
    report("main/side " + f, got, want[f])
if quick:
    sys.exit(0)
# 3. worker threads, side stream, through the loader's own preprocess
Loader = get_dataloader_by_name("stroke3-distributed")
hps = Loader.parse_hparams("max_seq_len=48")
for rep in range(3):
    ld = Loader(hps, d, seed=11, tokenizer=Tokenizer(centroids=cents))
    res = {}

    def work(f):
        res[f] = ld.preprocess(chunks[f])
    ths = [threading.Thread(target=work, args=(f,)) for f in chunks]
    [t.start() for t in ths]
    [t.join() for t in ths]
    n = sum(report(f"thread rep{rep} " + f, res[f], want[f]) for f in chunks)
    c = ld.tokenizer._dev_centroids
    print("  device table equals host table:", bool(np.array_equal(c.cpu().numpy(), cents)), flush=True)
# 4. the loader's iterator against the host loader
dev = Loader(hps, d, seed=11, tokenizer=Tokenizer(centroids=cents))
host = Loader(hps, d, seed=11, tokenizer=Tokenizer(centroids=cents), device=None)
a, b = dev.batch_iterator("train", 16), host.batch_iterator("train", 16)
for i in range(40):
    (xa, ya), (xb, yb) = next(a), next(b)
    if not np.array_equal(xa, xb):
        print("iterator batch", i, "rows differ:", np.flatnonzero((xa != xb).any(axis=1)).tolist(), flush=True)
print("iterator done; device table ok:", bool(np.array_equal(dev.tokenizer._dev_centroids.cpu().numpy(), cents)))

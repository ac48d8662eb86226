"""stroke-3 <-> token ids, the dictionary tokenizer of the reference (utils/tokenizer.py:16-101).

Vocabulary: PAD = 0, cluster id + 1 for each point (nearest KMeans centre of (dx, dy)), SEP = K + 1 after
every pen-lift point, SOS = K + 2, EOS = K + 3 (K = 1000 for the shipped dictionary => VOCAB_SIZE 1004).

The reference unpickles a scikit-learn 0.21 `KMeans` (prep_data/sketch_token/token_dict.pkl) and calls
`.predict`; that class path no longer exists in current scikit-learn and unpickling is arbitrary code
execution, so `load_centroids()` walks the pickle's opcode stream with `pickletools.genops` (which never
executes anything) and lifts out the `cluster_centers_` array bytes.  Encoding on the device
(`encode_batch_device`) goes through skf_tokenize_strokes in libskf.so; `encode` is the host reference
path used for single sketches and by the tests.
"""
import pickletools

import numpy as np


def load_centroids(dict_path):
    """(K, 2) float32 centres of the pickled sklearn KMeans, without unpickling it."""
    with open(dict_path, "rb") as f:
        data = f.read()
    seen_key, shape, dtype, want_shape = False, None, None, False
    ints = []
    for op, arg, _ in pickletools.genops(data):
        if op.name.endswith("UNICODE") and arg == "cluster_centers_":
            seen_key = True
            continue
        if not seen_key:
            continue
        if op.name == "MARK":
            ints = []
        elif op.name.startswith("BININT") or op.name in ("INT", "LONG1"):
            ints.append(int(arg))
        elif op.name.startswith("TUPLE") and shape is None and len(ints) >= 2:
            shape = tuple(ints[-2:])            # (n_clusters, 2)
        elif op.name.endswith("UNICODE") and isinstance(arg, str) and arg[:1] in "fd<>=|" and arg[-1:].isdigit():
            dtype = arg
        elif "BYTES" in op.name and shape is not None and len(arg) >= 8 * shape[0]:
            dt = np.dtype(dtype if dtype and dtype[0] in "<>=|" else "<" + (dtype or "f4"))
            arr = np.frombuffer(arg, dtype=dt).reshape(shape)
            return np.ascontiguousarray(arr, dtype=np.float32)
    raise ValueError(f"{dict_path}: no cluster_centers_ array found in the pickle stream")


class Tokenizer:
    def __init__(self, dict_path=None, max_seq_len=0, centroids=None):
        """utils/tokenizer.py:21-33.  `centroids` (K, 2) may be given directly instead of a pickle path."""
        self.max_seq_len = max_seq_len
        self.centroids = np.asarray(centroids, np.float32) if centroids is not None else load_centroids(dict_path)
        assert self.centroids.ndim == 2 and self.centroids.shape[1] == 2
        self.n_clusters = int(self.centroids.shape[0])
        self.PAD = 0
        self.SEP = self.n_clusters + 1
        self.SOS = self.n_clusters + 2
        self.EOS = self.n_clusters + 3
        self.VOCAB_SIZE = self.n_clusters + 4
        self._dev_centroids = None

    # ---- host path -------------------------------------------------------------------------
    def predict(self, xy):
        """KMeans.predict: index of the nearest centre (squared Euclidean distance, lowest index on ties)."""
        xy = np.asarray(xy, np.float32)
        out = np.empty(len(xy), np.int64)
        c = self.centroids
        for i in range(0, len(xy), 4096):
            blk = xy[i:i + 4096]
            dx = blk[:, None, 0] - c[None, :, 0]
            dy = blk[:, None, 1] - c[None, :, 1]
            out[i:i + 4096] = np.argmin(dx * dx + dy * dy, axis=1)
        return out

    def encode(self, stroke3, seq_len=0):
        """utils/tokenizer.py:35-64."""
        stroke3 = np.asarray(stroke3, np.float32).reshape(-1, 3)
        ids = self.predict(stroke3[:, :2]) + 1
        lift = stroke3[:, 2] == 1
        n = len(ids)
        pos = np.arange(n) + np.concatenate(([0], np.cumsum(lift)[:-1])) if n else np.zeros(0, np.int64)
        body = np.full(n + int(lift.sum()), self.SEP, np.int64)
        body[pos.astype(np.int64)] = ids
        out = np.concatenate(([self.SOS], body, [self.EOS]))
        if self.max_seq_len:
            npad = self.max_seq_len - len(out)
            if npad > 0:
                out = np.concatenate((out, np.full(npad, self.PAD, np.int64)))
            else:
                out = out[:self.max_seq_len]
                out[-2:] = [self.SEP, self.EOS]
        if len(out) < seq_len:
            out = np.concatenate((out, np.full(seq_len - len(out), self.PAD, np.int64)))
        return out

    def decode(self, seqs):
        """utils/tokenizer.py:66-101."""
        if len(seqs) > 0 and isinstance(seqs[0], (list, tuple, np.ndarray)):
            return [self.decode_single(np.squeeze(s)) for s in seqs]
        return self.decode_single(seqs)

    def decode_single(self, seq):
        ids, pens = [], []
        for i in np.asarray(seq).reshape(-1).tolist():
            if i not in (self.SOS, self.EOS, self.SEP, self.PAD):
                ids.append(i)
                pens.append(0)
            elif i == self.SEP and pens:
                pens[-1] = 1
            elif i == self.EOS:
                break
        if not ids:
            return np.zeros((1, 3), dtype=np.float32)
        dxy = self.centroids[np.asarray(ids) - 1]
        return np.c_[dxy, np.asarray(pens, np.float32)]

    # ---- device path -----------------------------------------------------------------------
    def prepare_device(self, device="cuda:0"):
        """Uploads the centre table once and waits for the device: the table outlives every stream that will read it,
        so its upload is not left to whichever prefetch thread happens to tokenise first."""
        import torch
        dev = torch.device(device)
        if self._dev_centroids is None or self._dev_centroids.device != dev:
            self._dev_centroids = torch.from_numpy(np.ascontiguousarray(self.centroids, np.float32)).to(dev).contiguous()
            torch.cuda.synchronize(dev)
        return self._dev_centroids

    def encode_batch_device(self, sketches, max_seq_len, device="cuda:0", limit=1000.0):
        """dataloaders/distributed_stroke3.py:90-125, 138-153 for a list of raw stroke-3 arrays, fused on the
        device: clip to +-limit, normalise by max(bbox_w, bbox_h, 1), nearest centre, SEP / SOS / EOS, truncate
        to max_seq_len, PAD.  Returns int64 [B, max_seq_len] in HBM (the train step reads it in place)."""
        import ctypes as C

        import torch

        from . import _lib
        lib = _lib.load()
        lens = np.array([len(s) for s in sketches], np.int64)
        offs = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
        flat = (np.concatenate([np.asarray(s, np.float32).reshape(-1, 3) for s in sketches]) if len(sketches)
                else np.zeros((0, 3), np.float32))
        dev = torch.device(device)
        self.prepare_device(dev)
        pts = torch.from_numpy(np.ascontiguousarray(flat)).to(dev)
        off_d = torch.from_numpy(offs).to(dev)
        out = torch.empty(len(sketches), max_seq_len, dtype=torch.int64, device=dev)
        _lib.check(lib.skf_tokenize_strokes(C.c_void_p(pts.data_ptr()), C.c_void_p(off_d.data_ptr()), len(sketches),
                                            C.c_void_p(self._dev_centroids.data_ptr()), self.n_clusters,
                                            C.c_float(limit), C.c_void_p(out.data_ptr()), max_seq_len,
                                            C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                   "skf_tokenize_strokes")
        return out


class GridTokenizer:
    """`token_type='grid'` (utils/tokenizer.py:103-200): a point of the ABSOLUTE polyline is the id of the grid cell
    it falls in, cells numbered row-major over [-1, 1)^2 with `resolution` cells per axis (+1: id 0 is PAD); SEP closes
    every stroke, SOS / EOS frame the sketch; padding / truncation as in the dictionary tokenizer.  Host path only
    (the reference builds it with resolution=100: vocabulary 10004).

    Kept quirks of the reference: coordinates are NOT clamped to the canvas (only the exact upper edge 1.0 is folded
    into the last cell; np.int64() truncates toward zero), and the points after the last pen lift are dropped unless
    the sketch has no pen lift at all (utils/skt_tools.py:71-96)."""

    def __init__(self, resolution=100, max_seq_len=0):
        self.max_seq_len = max_seq_len
        self.r = int(resolution / 2)
        self.resolution = 2 * self.r
        self.half_pixel = 1 / self.resolution
        self.PAD = 0
        self.SEP = self.resolution ** 2 + 1
        self.SOS = self.SEP + 1
        self.EOS = self.SEP + 2
        self.VOCAB_SIZE = self.resolution ** 2 + 4

    def _cells(self, v):
        c = np.int64((v + 1) * self.r)
        c[c == self.resolution] = self.resolution - 1
        return c

    def encode(self, stroke3, seq_len=0):
        s = np.asarray(stroke3, np.float64).reshape(-1, 3)
        xy = np.cumsum(s[:, :2], axis=0)                     # running sum in float64, point by point
        ends = np.flatnonzero(s[:, 2] == 1)
        bounds = [(0, len(s))] if len(ends) == 0 else list(zip(np.concatenate(([0], ends[:-1] + 1)), ends + 1))
        out = [self.SOS]
        for a, b in bounds:
            ids = self._cells(xy[a:b, 0]) + self._cells(xy[a:b, 1]) * self.resolution + 1
            out.extend(ids.tolist())
            out.append(self.SEP)
        out.append(self.EOS)
        if self.max_seq_len:
            npad = self.max_seq_len - len(out)
            if npad > 0:
                out += [self.PAD] * npad
            else:
                out = out[:self.max_seq_len]
                out[-2:] = [self.SEP, self.EOS]
        if len(out) < seq_len:
            out += [self.PAD] * (seq_len - len(out))
        return np.array(out, dtype=np.int64)

    def decode(self, seqs):
        if len(seqs) > 0 and isinstance(seqs[0], (list, tuple, np.ndarray)):
            return [self.decode_single(np.squeeze(s)) for s in seqs]
        return self.decode_single(seqs)

    def decode_single(self, tokens):
        """tokens -> stroke-3 whose first row is the absolute position of the first cell centre
        (lines_to_strokes(..., omit_first_point=False), utils/skt_tools.py:55-68)."""
        pts, pen = [], []
        for t in np.asarray(tokens).reshape(-1).tolist():
            if 0 < t < self.SEP:
                cy, cx = divmod(t - 1, self.resolution)
                pts.append((cx / self.r - 1 + self.half_pixel, cy / self.r - 1 + self.half_pixel))
                pen.append(0)
            elif t == self.SEP and pen and pen[-1] == 0:
                pen[-1] = 1
            elif t == self.EOS:
                break
        if not pts:
            pts, pen = [(0.0, 0.0)], [0]
        pen[-1] = 1                                          # the last point of the last line always closes it
        out = np.zeros((len(pts), 3), np.float64)
        out[:, :2] = np.asarray(pts, np.float64)
        out[:, 2] = pen
        out[1:, :2] -= out[:-1, :2].copy()
        return out


def shuffle_strokes(sketch, rng=None, perm=None):
    """`shuffle_stroke=True` (distributed_stroke3.py:107-110 with utils/skt_tools.py:55-96): the sketch as absolute
    polylines starting from the origin, the polylines in a new order (`perm`, or a permutation drawn from `rng`), back
    to stroke-3 deltas with the first point (the origin of the new first line) dropped.  float64, like the reference's
    Python-float accumulation; the points after the last pen lift are lost unless the sketch has no pen lift."""
    s = np.asarray(sketch).reshape(-1, 3)
    xy = np.cumsum(s[:, :2].astype(np.float64), axis=0)
    ends = np.flatnonzero(s[:, 2] == 1)
    if len(ends) == 0:
        lines = [np.vstack(([[0.0, 0.0]], xy))]
    else:
        starts = np.concatenate(([0], ends[:-1] + 1))
        lines = [xy[a:b + 1] for a, b in zip(starts, ends)]
        lines[0] = np.vstack(([[0.0, 0.0]], lines[0]))
    if perm is None:
        perm = (rng or np.random.default_rng()).permutation(len(lines))
    lines = [lines[i] for i in perm]
    out = np.zeros((sum(len(ln) for ln in lines), 3), np.float64)
    out[:, :2] = np.concatenate(lines)
    out[np.cumsum([len(ln) for ln in lines]) - 1, 2] = 1
    out[1:, :2] -= out[:-1, :2].copy()
    return out[1:]


def convert_to_absolute(sketch):
    """`use_absolute_strokes=True` (utils/sketch.py:169-175): running sum of the offsets in the array's own dtype,
    pen column unchanged."""
    s = np.array(sketch)
    s[:, :2] = np.cumsum(s[:, :2], axis=0, dtype=s.dtype)
    return s


def normalise_sketch(sketch, limit=1000.0, shuffle_rng=None, absolute=False, shuffle=False):
    """The part of DistributedStroke3DataLoader.preprocess shared by token and continuous mode
    (distributed_stroke3.py:93-112): clip to +-limit, float32, divide by max(bbox_w, bbox_h, 1) of the absolute path
    (origin included, utils/sketch.py:31-50), then the optional stroke shuffle and absolute conversion."""
    s = np.array(np.maximum(np.minimum(sketch, limit), -limit), dtype=np.float32).reshape(-1, 3)
    ax = np.cumsum(s[:, 0].astype(np.float64))
    ay = np.cumsum(s[:, 1].astype(np.float64))
    min_x, max_x = min(0.0, ax.min(initial=0.0)), max(0.0, ax.max(initial=0.0))
    min_y, max_y = min(0.0, ay.min(initial=0.0)), max(0.0, ay.max(initial=0.0))
    s[:, :2] /= max(max_x - min_x, max_y - min_y, 1)
    if shuffle:
        s = shuffle_strokes(s, shuffle_rng)
    if absolute:
        s = convert_to_absolute(s)
    return s


def preprocess_sketch_host(sketch, tokenizer, max_seq_len, limit=1000.0, shuffle=False, absolute=False, rng=None):
    """Host restatement of DistributedStroke3DataLoader.preprocess for one token-mode sketch
    (distributed_stroke3.py:90-125 + _cap_pad_and_convert_sketch 138-153, utils/sketch.py:31-50)."""
    s = normalise_sketch(sketch, limit, rng, absolute, shuffle)
    tok = tokenizer.encode(s)
    tok = tok[:max_seq_len]
    out = np.full(max_seq_len, tokenizer.PAD, np.int64)
    out[:len(tok)] = tok
    return out

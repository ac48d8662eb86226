"""Loader side of the drop-in boundary: the attribute surface the model plugin reads
from its `dataset` (dataloaders/distributed_stroke3.py:28-46, core/data.py:33-64, 146-172) plus a
synthetic loader of that shape used by bench.py / the tests (there is no dataset in this image).
"""
import numpy as np

from . import synth
from .hparams import HParams


class TokenizerInfo:
    """utils/tokenizer.py:29-33 with the 1000-centroid dictionary: the ids the model needs."""
    PAD, SEP, SOS, EOS, VOCAB_SIZE = synth.PAD, synth.SEP, synth.SOS, synth.EOS, synth.VOCAB_SIZE


class SyntheticStroke3DataLoader:
    """Same hparams / attributes / batch_iterator contract as `stroke3-distributed`, fed by
    sketchformer_b200.synth instead of .npz chunk files."""
    name = "stroke3-synthetic"

    @classmethod
    def default_hparams(cls):
        """dataloaders/distributed_stroke3.py:13-25"""
        return HParams(max_seq_len=200, shuffle_stroke=False, token_type="dictionary", use_continuous_data=False,
                       use_absolute_strokes=False, tokenizer_dict_file="prep_data/sketch_token/token_dict.pkl",
                       tokenizer_resolution=100, augment_stroke_prob=0.1, random_scale_factor=0.1)

    @classmethod
    def parse_hparams(cls, params):
        """core/data.py:54-64"""
        hps = cls.default_hparams()
        if params is not None:
            hps = hps.parse(params)
        return hps

    def __init__(self, hps, data_directory=None, n_classes=345, n_samples=1 << 20, seed=1234, rank=0, world_size=1,
                 vocab=synth.VOCAB_SIZE):
        self.hps = hps if isinstance(hps, dict) else dict(hps.values())
        self.data_directory = data_directory
        self.tokenizer = None if self.hps["use_continuous_data"] else TokenizerInfo()
        self.n_classes, self.n_samples = n_classes, n_samples
        self.class_names = np.array([f"class_{i}" for i in range(n_classes)])
        self.seed, self.rank, self.world_size, self.vocab = seed, rank, world_size, vocab

    def batch_iterator(self, split_name, batch_size, stop_at_end_of_split=False):
        """core/data.py:146-172: yields [x (B,T) int64 | (B,T,5) float32, y (B,1) int64].  With
        world_size > 1 every rank draws the same global batch and keeps rows [rank*B, (rank+1)*B)
        (SURVEY.md §8e "strong" partitioning)."""
        T, step = self.hps["max_seq_len"], 0
        gen = synth.continuous_batch if self.hps["use_continuous_data"] else synth.token_batch
        total = self.n_samples // (batch_size * self.world_size)
        while not (stop_at_end_of_split and step >= total):
            kw = {} if self.hps["use_continuous_data"] else {"vocab": self.vocab}
            x, y = gen(batch_size * self.world_size, T, self.n_classes, self.seed + step, **kw)
            lo = self.rank * batch_size
            yield [x[lo:lo + batch_size], y[lo:lo + batch_size]]
            step += 1


def augment_strokes(strokes, prob, rng):
    """utils/sketch.py:127-149: drop interior points of a stroke with probability `prob`, merging their
    offsets into the next kept point (points next to a pen lift are never dropped)."""
    result, prev, stroke, count = [], [0, 0, 1], [0, 0, 1], 0
    for i in range(len(strokes)):
        cand = [strokes[i][0], strokes[i][1], strokes[i][2]]
        count = 0 if (cand[2] == 1 or prev[2] == 1) else count + 1
        urnd = rng.random()
        if cand[2] == 0 and prev[2] == 0 and count > 2 and urnd < prob:
            stroke[0] += cand[0]
            stroke[1] += cand[1]
        else:
            stroke = cand
            prev = stroke
            result.append(stroke)
    return np.array(result, dtype=np.float32).reshape(-1, 3)


class _Split:
    """core/data.py:9-20 DatasetSplit."""

    def __init__(self, name, filepaths):
        self.name, self.filepaths, self.n_files = name, sorted(filepaths), len(filepaths)
        self.current = self.next = None
        self.current_index_shuffle, self.current_len, self.cursor = [], 0, 0
        self.file_cursor, self.file_shuffle, self.thread = 0, [], None
        self.did_load_first_megabatch = False


class DistributedStroke3DataLoader:
    """`stroke3-distributed` (dataloaders/distributed_stroke3.py:10-204 + core/data.py:33-254): the dataset is
    a directory of .npz chunks (`meta*.npz`, `train*.npz`, `valid*.npz`, `test*.npz`, written by
    prep_data/prepare-distributed-quickdraw.py:92-126) with `x` an object array of (n_i, 3) stroke-3 sketches
    and `y` integer labels.  One prefetch thread per split loads and preprocesses the next chunk while the
    current one is iterated; chunks are swapped when exhausted.

    Data parallel (not in the reference): with world_size > 1 the train chunks are dealt
    round-robin to the ranks (rank r reads files r, r + W, ...) when there are at least W of them, otherwise
    every rank walks the same chunk in the same (seeded) order and keeps samples r, r + W, ...
    """
    name = "stroke3-distributed"

    @classmethod
    def default_hparams(cls):
        return SyntheticStroke3DataLoader.default_hparams()

    @classmethod
    def parse_hparams(cls, params):
        hps = cls.default_hparams()
        if params is not None:
            hps = hps.parse(params)
        return hps

    def __init__(self, hps, data_directory, rank=0, world_size=1, seed=None, tokenizer=None, device="auto",
                 device_batches=False):
        """device: where whole chunks are preprocessed in dictionary-token mode — "auto" = the current CUDA device when
        there is one (skf_tokenize_strokes, one launch per chunk: the reference's per-sketch Python loop,
        distributed_stroke3.py:90-125, caps a 15 k sketches/s train step), None = the host restatement.
        device_batches: yield batches as CUDA tensors cut from the device-resident chunk (no per-step host->device
        copy); the default yields numpy arrays exactly like the reference."""
        import glob
        import os
        import threading
        self._threading = threading
        self.hps = hps if isinstance(hps, dict) else dict(hps.values())
        self.data_directory = data_directory
        self.rank, self.world_size = rank, world_size
        self.limit = 1000
        self.rng = np.random.default_rng(seed)
        # sample-slicing mode needs the SAME file order on every rank: a rank-independent generator for it
        self._file_rng = np.random.default_rng(0x5EED if seed is None else seed)
        self.tokenizer = tokenizer
        self.device, self.device_batches, self._loader_stream = None, bool(device_batches), None
        if device == "auto":
            try:
                import torch
                device = f"cuda:{torch.cuda.current_device()}" if torch.cuda.is_available() else None
            except Exception:      # noqa: BLE001
                device = None
        self.device = device
        # the prefetch threads of the three splits share one stream and one centre table on the device: chunks are
        # tokenised one at a time (tensors allocated, used and freed on that stream only)
        self._device_lock = threading.Lock()
        if not self.hps["use_continuous_data"] and tokenizer is None:      # distributed_stroke3.py:32-36
            from .tokenizer import GridTokenizer, Tokenizer
            if self.hps["token_type"] == "dictionary":
                self.tokenizer = Tokenizer(self.hps["tokenizer_dict_file"], max_seq_len=0)
            elif self.hps["token_type"] == "grid":
                self.tokenizer = GridTokenizer(resolution=100)            # the reference ignores tokenizer_resolution
            else:
                raise ValueError("token_type=%s: 'dictionary' or 'grid'" % self.hps["token_type"])
        files = glob.glob(os.path.join(data_directory, "*"))
        base = os.path.basename
        meta = [f for f in files if base(f).startswith("meta")][0]
        md = np.load(meta, allow_pickle=True)
        self.n_classes = int(md["n_classes"])
        self.n_samples = int(md["n_samples_train"])
        self.class_names = md["class_names"]
        self.scale_factor = float(md["std"])
        self.splits = {n: _Split(n, [f for f in files if base(f).startswith(n)]) for n in ("train", "test", "valid")}
        if self._device_path():
            # before any prefetch thread exists: the centre table is uploaded and the tokeniser kernel has run once
            # (module load, function attributes) on this thread, with the device idle afterwards
            import torch
            with torch.cuda.device(self.device):
                self.tokenizer.prepare_device(self.device)
                self.tokenizer.encode_batch_device([np.zeros((1, 3), np.float32)], int(self.hps["max_seq_len"]),
                                                   self.device, float(self.limit))
                torch.cuda.synchronize()
        tr = self.splits["train"]
        self._slice_samples = world_size > 1 and tr.n_files < world_size
        if world_size > 1 and not self._slice_samples:
            tr.filepaths = tr.filepaths[rank::world_size]
            tr.n_files = len(tr.filepaths)
        for s, split in self.splits.items():
            if not split.n_files:
                continue
            split.file_shuffle = self.reshuffle_file_indices(s, split.filepaths)
            self._start_load(split, split.filepaths[split.file_shuffle[0]])

    # ---- chunk loading (core/data.py:44-50, 79-113; distributed_stroke3.py:77-88) -----------------
    def _start_load(self, split, path):
        split.thread = self._threading.Thread(target=self.load_next_megabatch, args=(split.name, path), daemon=True)
        split.thread.start()

    def reshuffle_file_indices(self, split_name, filenames):
        if split_name != "train":
            return list(range(len(filenames)))
        # every rank walks the same chunk when the ranks slice samples instead of files
        return (self._file_rng if getattr(self, "_slice_samples", False) else self.rng).permutation(len(filenames))

    def reshuffle_sample_indices(self, split_name, data):
        if split_name != "train":
            return list(range(len(data["x"])))
        if self._slice_samples:      # every rank draws the same order, then keeps its residue class
            order = np.random.default_rng(len(data["x"]) + 7919 * self.splits["train"].file_cursor).permutation(len(data["x"]))
            return order[self.rank::self.world_size]
        return self.rng.permutation(len(data["x"]))

    def load_next_megabatch(self, split_name, selected_file):
        loaded = np.load(selected_file, allow_pickle=True)
        data = {"x": self.preprocess(loaded["x"], augment=split_name == "train"), "y": np.asarray(loaded["y"])}
        self.splits[split_name].next = data

    def _device_path(self):
        """Whole-chunk preprocessing on the device: dictionary tokens, relative strokes in file order (the options
        that reorder or re-reference the points, and the grid tokenizer, stay on the host restatement)."""
        return (self.device is not None and not self.hps["use_continuous_data"] and
                not self.hps["shuffle_stroke"] and not self.hps["use_absolute_strokes"] and
                hasattr(self.tokenizer, "encode_batch_device"))

    def _swap(self, split_name):
        split = self.splits[split_name]
        split.thread.join()
        split.current = split.next
        split.current_index_shuffle = self.reshuffle_sample_indices(split_name, split.current)
        split.current_len = len(split.current_index_shuffle)
        split.file_cursor += 1
        done = split.file_cursor == split.n_files
        if done:
            split.file_cursor = 0
            split.file_shuffle = self.reshuffle_file_indices(split_name, split.filepaths)
        if split.n_files > 1:
            self._start_load(split, split.filepaths[split.file_shuffle[split.file_cursor]])
        return done

    def _ready(self, split_name):
        split = self.splits[split_name]
        if not split.did_load_first_megabatch:
            split.did_load_first_megabatch = True
            self._swap(split_name)
        return split

    # ---- per-sketch preprocessing (distributed_stroke3.py:90-162) ---------------------------------
    def preprocess(self, data, augment=False):
        """distributed_stroke3.py:90-125: clip, (continuous + train: random scale and point dropout), bbox
        normalisation, optional stroke shuffle / absolute coordinates, tokenizer, cut to max_seq_len, PAD / stroke-5."""
        from .tokenizer import convert_to_absolute, preprocess_sketch_host, shuffle_strokes
        if self._device_path() and len(data):
            # one kernel launch for the whole chunk, on the loader's own stream (this runs in the prefetch thread while
            # the train step owns the device's other streams); bit-identical to the host loop below
            import torch
            with self._device_lock, torch.cuda.device(self.device):
                if self._loader_stream is None:
                    self._loader_stream = torch.cuda.Stream(self.device)
                with torch.cuda.stream(self._loader_stream):
                    tok = self.preprocess_on_device(list(data), self.device)
                    out = tok if self.device_batches else tok.cpu().numpy()
                    self._loader_stream.synchronize()
                if self.device_batches:
                    # the chunk outlives this stream context: the train step reads it from its own stream
                    out.record_stream(torch.cuda.current_stream(self.device))
            return out
        T, out = self.hps["max_seq_len"], []
        shuffle, absolute = bool(self.hps["shuffle_stroke"]), bool(self.hps["use_absolute_strokes"])
        for sketch in data:
            if not self.hps["use_continuous_data"]:
                out.append(preprocess_sketch_host(sketch, self.tokenizer, T, self.limit, shuffle, absolute, self.rng))
                continue
            s = np.array(np.maximum(np.minimum(sketch, self.limit), -self.limit), dtype=np.float32).reshape(-1, 3)
            if augment and self.hps["augment_stroke_prob"] > 0:
                e = self.hps["random_scale_factor"]
                s = s.copy()
                s[:, 0] *= (self.rng.random() - 0.5) * 2 * e + 1.0
                s[:, 1] *= (self.rng.random() - 0.5) * 2 * e + 1.0
                s = augment_strokes(s, self.hps["augment_stroke_prob"], self.rng)
            ax, ay = np.cumsum(s[:, 0].astype(np.float64)), np.cumsum(s[:, 1].astype(np.float64))
            w = max(0.0, ax.max(initial=0.0)) - min(0.0, ax.min(initial=0.0))
            h = max(0.0, ay.max(initial=0.0)) - min(0.0, ay.min(initial=0.0))
            s[:, :2] /= max(w, h, 1)
            if shuffle:
                s = shuffle_strokes(s, self.rng)
            if absolute:
                s = convert_to_absolute(s)
            s = s[:T]
            conv = np.zeros((T, 5), dtype=np.float32)
            n = len(s)
            conv[:n, 0:2] = s[:, 0:2]
            conv[:n, 3] = s[:, 2]
            conv[:n, 2] = 1 - s[:, 2]
            conv[n:, 4] = 1
            conv[-1:, 4] = 1
            out.append(conv)
        return np.array(out)

    def preprocess_on_device(self, sketches, device="cuda:0"):
        """Token mode: the same preprocessing as `preprocess`, fused in one CUDA kernel (skf_tokenize_strokes);
        returns int64 [B, max_seq_len] already in HBM."""
        if not hasattr(self.tokenizer, "encode_batch_device"):
            raise NotImplementedError("device-side tokenisation exists for the dictionary tokenizer only "
                                      "(token_type=grid is tokenised on the host: preprocess())")
        return self.tokenizer.encode_batch_device(sketches, self.hps["max_seq_len"], device, float(self.limit))

    # ---- iteration (core/data.py:115-254) -----------------------------------------------------------
    def get_sample(self, data, idx):
        return [np.array(data["x"][idx]), np.expand_dims(data["y"][idx], axis=-1)]

    def _take(self, split, indices):
        """Samples `indices` of the loaded chunk as one stacked batch (fancy indexing: no per-sample Python work)."""
        idx = np.asarray(indices, dtype=np.int64)
        x = split.current["x"]
        if isinstance(x, np.ndarray):
            bx = x[idx]
        else:                      # device-resident chunk (device_batches=True)
            import torch
            bx = x.index_select(0, torch.as_tensor(idx, device=x.device))
        by = np.asarray(split.current["y"])[idx].reshape(len(idx), -1)[:, :1].astype(np.int64)
        return bx, by

    def _collate(self, split, indices):
        bx, by = self._take(split, indices)
        return [bx, by]

    def batch_iterator(self, split_name, batch_size, stop_at_end_of_split=False):
        """core/data.py:115-172 (sample_iterator + batch_iterator): a batch continues across a chunk swap; when
        the last file of the split is exhausted the running batch is cut short ('end_of_split')."""
        split = self._ready(split_name)
        status = "running"
        while not stop_at_end_of_split or status != "end_of_split":
            # The reference draws one sample at a time (core/data.py:115-172); the same sequence in slices: take what
            # the loaded chunk still has, swap chunks when it is exhausted, and when that swap wrapped the split the
            # batch ends with the single sample drawn right after it.
            xs, ys, need, status = [], [], batch_size, "running"
            while status == "running":
                if split.cursor >= split.current_len:
                    end = self._swap(split_name)
                    split.cursor = 0
                    if end:
                        status, need = "end_of_split", 1
                take = min(need, split.current_len - split.cursor)
                bx, by = self._take(split, split.current_index_shuffle[split.cursor:split.cursor + take])
                split.cursor += take
                need -= take
                xs.append(bx)
                ys.append(by)
                if status == "running" and need == 0:
                    status = "end_of_batch"
            if len(xs) == 1:
                yield [xs[0], ys[0]]
            elif isinstance(xs[0], np.ndarray):
                yield [np.concatenate(xs), np.concatenate(ys)]
            else:
                import torch
                yield [torch.cat(xs), np.concatenate(ys)]

    def get_n_samples_from(self, split_name, n, shuffled=False, seeded=None, preprocess=False):
        """core/data.py:174-209: the first n samples of the loaded chunk in FILE order, or n of a permutation of it.
        `seeded` seeds numpy's legacy global generator in the reference (np.random.seed(int(seeded)) followed by
        np.random.permutation), so the same MT19937 stream is used here: seeded=True picks the very samples the
        reference's evaluation picks (models/evaluation_mixin.py:25-27)."""
        split = self._ready(split_name)
        if shuffled:
            perm = (np.random.RandomState(int(seeded)).permutation(split.current_len) if seeded is not None
                    else self.rng.permutation(split.current_len))
            order = perm[:n]
        else:
            order = list(range(min(n, split.current_len)))
        return self._collate(split, order)

    def get_all_data_from(self, split_name):
        split = self._ready(split_name)
        return self.get_n_samples_from(split_name, split.current_len)


dataloaders_by_name = {SyntheticStroke3DataLoader.name: SyntheticStroke3DataLoader,
                       DistributedStroke3DataLoader.name: DistributedStroke3DataLoader}


def get_dataloader_by_name(name):
    """dataloaders/__init__.py:16-17"""
    return dataloaders_by_name[name]

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

    def __init__(self, hps, data_directory=None, n_classes=345, n_samples=1 << 20, seed=1234, rank=0, world_size=1):
        self.hps = hps if isinstance(hps, dict) else dict(hps.values())
        self.data_directory = data_directory
        self.tokenizer = None if self.hps["use_continuous_data"] else TokenizerInfo()
        self.n_classes, self.n_samples = n_classes, n_samples
        self.class_names = np.array([f"class_{i}" for i in range(n_classes)])
        self.seed, self.rank, self.world_size = seed, rank, world_size

    def batch_iterator(self, split_name, batch_size, stop_at_end_of_split=False):
        """core/data.py:146-172: yields [x (B,T) int64 | (B,T,5) float32, y (B,1) int64].  With
        world_size > 1 every rank draws the same global batch and keeps rows [rank*B, (rank+1)*B)
        (SURVEY.md §8e "strong" partitioning)."""
        T, step = self.hps["max_seq_len"], 0
        gen = synth.continuous_batch if self.hps["use_continuous_data"] else synth.token_batch
        total = self.n_samples // (batch_size * self.world_size)
        while not (stop_at_end_of_split and step >= total):
            x, y = gen(batch_size * self.world_size, T, self.n_classes, self.seed + step)
            lo = self.rank * batch_size
            yield [x[lo:lo + batch_size], y[lo:lo + batch_size]]
            step += 1


dataloaders_by_name = {SyntheticStroke3DataLoader.name: SyntheticStroke3DataLoader}


def get_dataloader_by_name(name):
    """dataloaders/__init__.py:16-17"""
    return dataloaders_by_name[name]

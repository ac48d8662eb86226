"""Synthetic stroke batches.  numpy only; used by bench.py, the tests and
the oracle fixtures so that every arm sees byte-identical inputs.

Token layout follows utils/tokenizer.py:29-64 + dataloaders/distributed_stroke3.py:114-153 of
the reference: [SOS] body [EOS] PAD*, body in 1..1000 with SEP=1001 after pen lifts.
"""
import numpy as np

PAD, SEP, SOS, EOS, VOCAB_SIZE = 0, 1001, 1002, 1003, 1004


def token_batch(batch, seq_len, n_classes, seed=1234, ragged=True, vocab=VOCAB_SIZE):
    """x int64 (B,T), y int64 (B,1).  ragged: n ~ U{T/4..T} total tokens, else n = T."""
    rng = np.random.default_rng(seed)
    sep, sos, eos = vocab - 3, vocab - 2, vocab - 1
    x = np.zeros((batch, seq_len), dtype=np.int64)
    for b in range(batch):
        n = int(rng.integers(max(3, seq_len // 4), seq_len + 1)) if ragged else seq_len
        body = rng.integers(1, vocab - 3, size=n - 2)
        lift = rng.random(n - 2) < 0.08
        body = np.where(lift, sep, body)
        x[b, 0] = sos
        x[b, 1:n - 1] = body
        x[b, n - 1] = eos
    y = rng.integers(0, n_classes, size=(batch, 1)).astype(np.int64)
    return x, y


def continuous_batch(batch, seq_len, n_classes, seed=1234):
    """stroke-5 float32 (B,T,5) rows [dx,dy,p_down,p_up,p_pad]; y int64 (B,1)."""
    rng = np.random.default_rng(seed)
    x = np.zeros((batch, seq_len, 5), dtype=np.float32)
    for b in range(batch):
        n = int(rng.integers(max(2, seq_len // 4), seq_len))
        x[b, :n, 0:2] = rng.normal(0.0, 0.05, size=(n, 2)).astype(np.float32)
        up = rng.random(n) < 0.08
        x[b, :n, 3] = up
        x[b, :n, 2] = 1 - x[b, :n, 3]
        x[b, n:, 4] = 1
        x[b, -1, 4] = 1
    y = rng.integers(0, n_classes, size=(batch, 1)).astype(np.int64)
    return x, y

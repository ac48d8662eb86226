"""TEST INFRASTRUCTURE (only tests/, smoke() and bench.py's CPU legs may import oracle/).

numpy mirror of the counter-based dropout mask of the CUDA path (skf_drop_apply in
sketchformer_b200/csrc/skf_common.cuh): the reference's tf.keras.layers.Dropout draws from TF's stateful
RNG, which no other implementation can reproduce, so parity under dropout is checked by giving the oracle
the SAME keep-mask the device generates: keep(i) = (fmix32(fmix32(i + GOLDEN*(site+1)) ^ key) >> 8) >= rate*2^24,
key = seed ^ (step * 0x632BE5AB), all in uint32 arithmetic.
"""
import numpy as np

SITE_ENC_X0, SITE_DEC_X0, SITE_ENC_L, SITE_DEC_L, SITE_CBUF = 1, 2, 10, 100, 1000


def _fmix32(h):
    h = h.astype(np.uint32)
    h ^= h >> np.uint32(16)
    h = (h * np.uint32(0x85EBCA6B)).astype(np.uint32)
    h ^= h >> np.uint32(13)
    h = (h * np.uint32(0xC2B2AE35)).astype(np.uint32)
    h ^= h >> np.uint32(16)
    return h


def keep_mask(shape, rate, seed, site, step):
    """bool array `shape`: True where the element is kept (element index = C-order linear index)."""
    n = int(np.prod(shape))
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64).astype(np.uint32)
        key = np.uint32(seed) ^ np.uint32((int(step) * 0x632BE5AB) & 0xFFFFFFFF)
        h = _fmix32(_fmix32(idx + np.uint32((0x9E3779B9 * (int(site) + 1)) & 0xFFFFFFFF)) ^ key)
    thresh = np.uint32(np.float32(rate) * np.float32(16777216.0))
    return ((h >> np.uint32(8)) >= thresh).reshape(shape)


def site_id(name):
    """Oracle dropout-site name -> device site id (skf_plan.cu)."""
    if name == "enc.x0":
        return SITE_ENC_X0
    if name == "dec.x0":
        return SITE_DEC_X0
    if name.startswith("cbuf"):
        return SITE_CBUF + int(name[4:])
    side, layer, kind = name.split(".")
    i = int(layer[1:])
    if side == "enc":
        return SITE_ENC_L + 2 * i + {"attn": 0, "ffn": 1}[kind]
    return SITE_DEC_L + 3 * i + {"attn1": 0, "attn2": 1, "ffn": 2}[kind]


class DeviceMasks:
    """Mask provider for oracle.dropout_masks(): the masks the CUDA plan draws at optimizer step `step`."""

    def __init__(self, rate, seed, step, class_rate=0.1):
        self.rate, self.seed, self.step, self.class_rate = rate, seed, step, class_rate

    def __call__(self, site, shape):
        rate = self.class_rate if site.startswith("cbuf") else self.rate
        return keep_mask(tuple(shape), rate, self.seed, site_id(site), self.step)

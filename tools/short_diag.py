"""GPU diagnostic: the whole-row attention kernels and the cfg2 train step on SHORT sequences (loader-like lengths
23..140 and tiny ones), many items per CTA.
   python tools/short_diag.py op            op level, backends 6 / 7 against the SIMT kernels
   python tools/short_diag.py eng [k=v ..]  Transformer.train_on_batch on short batches (options via skf_set_option)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from sketchformer_b200 import _lib  # noqa: E402
from sketchformer_b200.ops import ptr, stream  # noqa: E402

lib = _lib.load()


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def op_level():
    B, H, dh = 128, 8, 32
    d = H * dh
    for (lo, hi) in ((23, 140), (2, 40), (100, 135)):
        for (T, mode) in ((200, 1), (199, 2)):
            g = torch.Generator(device="cuda").manual_seed(lo + T)
            qkv = torch.randn(B * T, 3 * d, device="cuda", generator=g)
            dO = torch.randn(B * T, d, device="cuda", generator=g)
            lens = torch.randint(lo, hi + 1, (B, 1), device="cuda", generator=g)
            pad = (torch.arange(T, device="cuda")[None, :] >= lens).to(torch.uint8).contiguous()
            res = {}
            for be in (1, 6, 7):
                O = torch.zeros(B * T, d, device="cuda"); lse = torch.zeros(B, H, T, device="cuda")
                delta = torch.zeros(B, H, T, device="cuda"); dqkv = torch.zeros_like(qkv)
                img = torch.empty(lib.skf_attn_image_bytes(B, H, T, T, 0), dtype=torch.uint8, device="cuda")
                img_do = torch.empty(lib.skf_attn_image_bytes(B, H, T, T, 1), dtype=torch.uint8, device="cuda")
                q, k, v = qkv, qkv[:, d:], qkv[:, 2 * d:]
                for rep in range(2):            # twice: the second run starts from the first one's barrier-free state
                    try:
                        if be == 7:
                            _lib.check(lib.skf_attn_fwd_img(ptr(q), 3 * d, ptr(k), 3 * d, ptr(v), 3 * d, ptr(O), d, ptr(lse),
                                                            ptr(pad), B, H, T, T, dh, mode, ptr(img), stream()))
                        else:
                            _lib.check(lib.skf_attn_fwd(ptr(q), 3 * d, ptr(k), 3 * d, ptr(v), 3 * d, ptr(O), d, ptr(lse),
                                                        ptr(pad), B, H, T, T, dh, mode, 0, be, stream()))
                        torch.cuda.synchronize()
                        print(f"lens {lo}-{hi} T={T} mode={mode} backend {be} fwd ok", flush=True)
                        if be == 7:
                            _lib.check(lib.skf_attn_bwd_img(ptr(q), 3 * d, ptr(k), 3 * d, ptr(v), 3 * d, ptr(O), d, ptr(dO), d,
                                                            ptr(lse), ptr(delta), ptr(dqkv), 3 * d, ptr(dqkv[:, d:]), 3 * d,
                                                            ptr(dqkv[:, 2 * d:]), 3 * d, ptr(pad), B, H, T, T, dh, mode, ptr(img),
                                                            ptr(img_do), stream()))
                        else:
                            _lib.check(lib.skf_attn_bwd(ptr(q), 3 * d, ptr(k), 3 * d, ptr(v), 3 * d, ptr(O), d, ptr(dO), d,
                                                        ptr(lse), ptr(delta), ptr(dqkv), 3 * d, ptr(dqkv[:, d:]), 3 * d,
                                                        ptr(dqkv[:, 2 * d:]), 3 * d, ptr(pad), B, H, T, T, dh, mode, 0, be, stream()))
                        torch.cuda.synchronize()
                        print(f"lens {lo}-{hi} T={T} mode={mode} backend {be} bwd ok", flush=True)
                    except Exception as e:      # noqa: BLE001
                        print(f"lens {lo}-{hi} T={T} mode={mode} backend {be} FAILED: {str(e)[:300]}", flush=True)
                        sys.exit(1)
                res[be] = (O.clone(), dqkv.clone())
            for be in (6, 7):
                print(f"   backend {be} vs SIMT: O {rel(res[be][0], res[1][0]):.2e}  dQKV {rel(res[be][1], res[1][1]):.2e}",
                      flush=True)


def short_batch(B, T, lo, hi, seed):
    rng = np.random.default_rng(seed)
    x = np.zeros((B, T), np.int64)
    for b in range(B):
        n = int(rng.integers(lo, hi + 1))
        body = rng.integers(1, 1001, size=n - 2)
        body = np.where(rng.random(n - 2) < 0.08, 1001, body)
        x[b, 0], x[b, 1:n - 1], x[b, n - 1] = 1002, body, 1003
    return x, rng.integers(0, 345, size=(B, 1)).astype(np.int64)


def eng_level(opts):
    from sketchformer_b200.data import get_dataloader_by_name
    from sketchformer_b200.model import get_model_by_name
    for kv in opts:
        k, v = kv.split("=")
        _lib.check(lib.skf_set_option(k.encode(), int(v)), "skf_set_option")
    Model = get_model_by_name("sketch-transformer-tf2")
    Loader = get_dataloader_by_name("stroke3-synthetic")
    dataset = Loader(Loader.parse_hparams("max_seq_len=200"), None, n_classes=345, seed=1)
    model = Model(Model.parse_hparams(base="batch_size=128", specific="num_layers=4,d_model=256,dff=1024,num_heads=8,"
                                      "lowerdim=256,dropout_rate=0.0"), dataset, None, "diag", device="cuda:0", seed=0)
    for i, (lo, hi) in enumerate([(50, 200), (50, 200), (23, 140), (23, 140), (23, 140), (3, 40), (3, 40), (100, 135), (23, 140)]):
        try:
            m = model.train_on_batch(short_batch(128, 200, lo, hi, i))
            torch.cuda.synchronize()
            print(f"step {i} lens {lo}-{hi} ok total_loss {m['total_loss']:.4f}", flush=True)
        except Exception as e:      # noqa: BLE001
            print(f"step {i} lens {lo}-{hi} FAILED {str(e)[:300]}", flush=True)
            sys.exit(1)


if __name__ == "__main__":
    if sys.argv[1] == "op":
        op_level()
    else:
        eng_level(sys.argv[2:])

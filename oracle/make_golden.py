"""Generate tests/golden/*.npz from the oracle (run here, committed with the vectors).

    python -m oracle.make_golden

These vectors freeze the *oracle's* outputs (TensorFlow cannot run in this image and the reference ships
no fixtures); they guard the restatement against drift and give the CUDA parity tests byte-identical
inputs/expected values on the GPU box, where /root/reference is absent.  The vectors made by the
REFERENCE's own code are written by make_reference_golden.py / make_tokenizer_golden.py /
make_loader_golden.py; against TensorFlow's kernels parity stays unpinned (see sketchformer_oracle.py).
"""
import os
import sys
import zlib

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sketchformer_oracle as O          # noqa: E402
from sketchformer_b200 import synth                   # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

TINY = dict(num_layers=2, d_model=64, dff=128, num_heads=4, lowerdim=32, max_seq_len=24,
            vocab_size=68, n_classes=10)
CFG1 = dict(num_layers=4, d_model=128, dff=512, num_heads=8, lowerdim=256, max_seq_len=64,
            vocab_size=1004, n_classes=10)


def crc(t):
    return np.uint32(zlib.crc32(np.ascontiguousarray(t.detach().numpy()).tobytes()))


def run(cfg_kw, batch, seed, continuous, full, steps=3):
    cfg = O.OracleConfig(use_continuous_data=continuous, **cfg_kw)
    out = {}
    if continuous:
        x, y = synth.continuous_batch(batch, cfg.max_seq_len, cfg.n_classes, seed)
    else:
        x, y = synth.token_batch(batch, cfg.max_seq_len, cfg.n_classes, seed, vocab=cfg.vocab_size)
    out["x"], out["y"] = x, y
    for dt_name, dt in (("f64", torch.float64), ("f32", torch.float32)):
        p = O.init_params(cfg, 0, dt)
        if dt_name == "f32":
            out["param_crc"] = np.array([crc(v) for v in p.values()], dtype=np.uint32)
        opt = O.AdamTF(p, cfg.d_model)
        xt = torch.from_numpy(x)
        yt = torch.from_numpy(y)
        for s in range(steps):
            tape = {} if s == 0 else None
            losses, grads = O.train_step(p, opt, xt, yt, cfg, tape=tape, dtype=dt)
            for k, v in losses.items():
                out[f"{dt_name}/step{s}/{k}"] = np.float64(v)
            if s == 0:
                for k, v in tape.items():
                    a = v.detach().to(torch.float64).numpy()
                    if dt_name == "f64":
                        out[f"act_norm/{k}"] = np.float64(np.linalg.norm(a))
                        if full:
                            out[f"act/{k}"] = a.astype(np.float32)
                        else:
                            out[f"act_head/{k}"] = a.reshape(-1)[:64].astype(np.float32)
                for k, v in grads.items():
                    a = v.to(torch.float64).numpy()
                    if dt_name == "f64":
                        out[f"grad_norm/{k}"] = np.float64(np.linalg.norm(a))
                        if full:
                            out[f"grad/{k}"] = a.astype(np.float32)
                        else:
                            out[f"grad_head/{k}"] = a.reshape(-1)[:64].astype(np.float32)
        if dt_name == "f64":
            for k, v in p.items():
                out[f"pnorm_after/{k}"] = np.float64(v.norm())
                if full:
                    out[f"p_after/{k}"] = v.numpy().astype(np.float32)
            # inference: class prediction + greedy decode from the trained-3-steps weights
            xi = torch.from_numpy(x)
            if cfg.do_classification:
                pc = O.predict_class(p, xi, cfg, dt)
                out["pred/embedding"] = pc["embedding"].numpy().astype(np.float32)
                out["pred/class"] = pc["class"].numpy()
            if full and cfg.do_classification:
                pr = O.predict(p, xi[:2], cfg, sos=cfg.vocab_size - 2, eos=cfg.vocab_size - 1, dtype=dt)
                out["pred/recon"] = pr["recon"].numpy()
    for k, v in cfg.to_dict().items():
        out["cfg/" + k] = np.array(v)
    return out


# hparam variants of models/sketchformer.py:24-53 that change the graph: SelfAttnV2 + one class buffer layer,
# no bottleneck (lowerdim=0: the decoder attends to the encoder output), decoder cross mask = encoder padding mask
TINY_V2 = dict(TINY, attn_version=2, class_buffer_layers=1, class_dropout=0.0)   # dropout off: replayable
TINY_LOWERDIM0 = dict(TINY, lowerdim=0, do_classification=False)
TINY_NOTBLIND = dict(TINY, blind_decoder_mask=False)
FIXTURES = [("tiny_tokens.npz", TINY, 4, False, True), ("tiny_continuous.npz", TINY, 4, True, True),
            ("cfg1_tokens.npz", CFG1, 8, False, False), ("tiny_v2.npz", TINY_V2, 4, False, False),
            ("tiny_lowerdim0.npz", TINY_LOWERDIM0, 4, False, False), ("tiny_notblind.npz", TINY_NOTBLIND, 4, False, False)]


def main(force=False):
    """Writes the fixtures that do not exist yet (all of them with --force: the zip container is not
    byte-reproducible, so existing files are left alone unless asked)."""
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    for name, kw, batch, cont, full in FIXTURES:
        path = os.path.join(GOLD, name)
        if force or not os.path.exists(path):
            np.savez_compressed(path, **run(kw, batch, 1234, cont, full))
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main(force="--force" in sys.argv)

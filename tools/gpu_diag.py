"""Verbose parity table (every activation / gradient) of the CUDA train step vs the oracle.
Run on the GPU box:  python tools/gpu_diag.py --cfg cfg1 --backend 0"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sketchformer_oracle as O            # noqa: E402
from oracle.make_golden import CFG1, TINY              # noqa: E402
from sketchformer_b200 import synth, _lib              # noqa: E402
from tests.gpu_util import rel_err                     # noqa: E402
from tests.test_gpu_train import build, engine_activations, engine_relu_pattern   # noqa: E402

CFG2 = dict(num_layers=4, d_model=256, dff=1024, num_heads=8, lowerdim=256, max_seq_len=200, vocab_size=1004,
            n_classes=345)

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="tiny")
    ap.add_argument("--backend", type=int, default=1)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--attn-backend", type=int, default=-1, help="-1: same as --backend")
    ap.add_argument("--tma-tf32", type=int, default=1)
    ap.add_argument("--precision", type=int, default=2)
    ap.add_argument("--quiet", action="store_true")
    ap.add_argument("--all", action="store_true")
    ap.add_argument("--free-relu", action="store_true", help="do not impose the engine's ReLU pattern on the oracle")
    a = ap.parse_args()
    kw, B = {"tiny": (TINY, 4), "cfg1": (CFG1, 8), "cfg2": (CFG2, 16)}[a.cfg]
    B = a.batch or B
    _lib.check(_lib.load().skf_set_option(b"tma_tf32", a.tma_tf32))
    ocfg, eng = build(kw, B, a.backend, a.precision, attn_backend=a.attn_backend)
    x, y = synth.token_batch(B, ocfg.max_seq_len, ocfg.n_classes, 1234, vocab=ocfg.vocab_size)
    xt, yt = torch.from_numpy(x), torch.from_numpy(y)
    p = O.init_params(ocfg, 0, torch.float64)
    eng.load_params(p)
    opt = O.AdamTF(p, ocfg.d_model)
    tape, own = {}, {}
    eng.train_step(xt.cuda(), yt.cuda(), True)
    torch.cuda.synchronize()
    pattern = engine_relu_pattern(eng, ocfg, B) if not a.free_relu else None
    t0 = time.time()
    with O.relu_pattern(pattern, record=own):
        losses, grads = O.train_step(p, opt, xt, yt, ocfg, tape=tape, dtype=torch.float64)
    print(f"oracle fp64 step: {time.time() - t0:.2f}s", losses)
    if pattern:
        print("relu flips:", sum(int((own[k] != pattern[k]).sum()) for k in pattern), "of",
              sum(v.numel() for v in pattern.values()))
    print("engine:", eng.last_step_metrics(), "launches", eng.launches())
    acts = engine_activations(eng, ocfg, B)
    worst = 0.0
    for k, ref in tape.items():
        if k not in acts:          # not materialised by the engine (folded cross attention: pre_decoder, kv2, ...)
            continue
        e = rel_err(acts[k], ref)
        worst = max(worst, e)
        if a.all or (e > 1e-4 and not a.quiet):
            print(f"  act  {k:28s} {e:.3e}")
    print(f"worst activation rel err {worst:.3e}")
    got = eng.export_grads(grads)
    G = float(torch.sqrt(sum((v.double() ** 2).sum() for v in grads.values())))
    worst = 0.0
    for k, ref in grads.items():
        num = float((got[k].double() - ref.double()).norm())
        e = num / max(float(ref.norm()), 1e-4 * G)
        worst = max(worst, e)
        if a.all or (e > 1e-4 and not a.quiet):
            print(f"  grad {k:28s} {e:.3e}   |ref|={float(ref.norm()):.3e}")
    top = sorted(((float((got[k].double() - grads[k].double()).norm()) / max(float(grads[k].norm()), 1e-4 * G), k)
                  for k in grads), reverse=True)[:6]
    print("top grad errs:", [(k, f"{e:.2e}") for e, k in top])
    print(f"worst gradient rel err {worst:.3e}  (global grad norm {G:.3e})")

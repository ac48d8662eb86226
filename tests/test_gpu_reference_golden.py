"""The CUDA engine against vectors made by the REFERENCE'S OWN model code (tests/golden/reference_model_tiny.npz,
written by oracle/make_reference_golden.py from /root/reference on oracle/tf_shim.py): embedding, class prediction,
both losses and the greedy token sequences for the graph-changing hparam variants.  The other GPU tests compare the
engine with the oracle and the CPU tests compare the oracle with these vectors; this file closes the chain directly."""
import os

import numpy as np
import pytest
import torch

from oracle import sketchformer_oracle as O
from oracle.make_reference_golden import VARIANTS, inputs, params

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_model_tiny.npz")


@pytest.mark.parametrize("name,kw", [v for v in VARIANTS if v[0] != "continuous"], ids=[v[0] for v in VARIANTS if v[0] != "continuous"])
@pytest.mark.parametrize("backend", [1, 0], ids=["simt", "tcgen05"])
def test_engine_matches_vectors_made_by_the_reference_model_code(name, kw, backend):
    from tests.gpu_util import rel_err
    from tests.test_gpu_train import build
    g = np.load(GOLDEN)
    tol = 2e-4 if backend == 1 else 1e-3
    cfg = O.OracleConfig(**kw)
    flags = {k: v for k, v in kw.items() if k in ("attn_version", "class_buffer_layers", "class_dropout", "lowerdim",
                                                   "do_classification", "blind_decoder_mask")}
    base = {k: v for k, v in kw.items() if k not in flags}
    B = 4
    ocfg, eng = build(dict(base, **({"lowerdim": flags.pop("lowerdim")} if "lowerdim" in flags else {})), B, backend, **flags)
    p = params(cfg)
    x, y = inputs(cfg)
    assert np.array_equal(x.numpy(), g[f"{name}/x"])
    eng.load_params(p)
    xd, yd = x.cuda(), y.cuda()
    # losses of the forward pass at these weights (lr(0) = 0: the first optimizer step leaves them unchanged)
    eng.train_step(xd, yd, apply_adam=True)
    torch.cuda.synchronize()
    got = eng.last_step_metrics()
    assert got["recon_loss"] == pytest.approx(float(g[f"{name}/recon_loss"]), rel=tol)
    if f"{name}/class_loss" in g:
        assert got["class_loss"] == pytest.approx(float(g[f"{name}/class_loss"]), rel=tol)
    # inference entry points
    E = g[f"{name}/embedding"].shape[-1]
    eng.encode(xd)
    if cfg.lowerdim:
        assert rel_err(eng.tensor("embedding", (B, E)), torch.from_numpy(g[f"{name}/embedding"])) < tol
    if f"{name}/pred_class" in g:
        assert eng.tensor("cls.argmax", dtype=torch.int32).cpu()[:B].tolist() == g[f"{name}/pred_class"].tolist()
    if cfg.blind_decoder_mask and cfg.lowerdim:
        from tests.test_gpu_decode import _assert_same_tokens
        tok = eng.decode(None, None, cfg.vocab_size - 2, cfg.vocab_size - 1).cpu().numpy()
        _assert_same_tokens(p, ocfg, torch.from_numpy(g[f"{name}/embedding"]), tok, g[f"{name}/greedy"])

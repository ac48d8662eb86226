"""Data-parallel host logic on CPU (gloo, world_size 2): the sharded loader hands every rank a disjoint slice
of the same global batch, and summing per-rank gradients with the 1/world factor that the Adam kernel folds in
(skf_adam_step grad_scale) reproduces the full-batch gradient of the oracle — the property the
NCCL all-reduce of the flat gradient arena relies on.  The arena layout and the gradient buckets are the library's own
(skf_param_table / skf_grad_buckets, host code); no GPU: the per-rank gradients come from the oracle."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sketchformer_oracle as O
from oracle.make_golden import TINY
from sketchformer_b200.data import get_dataloader_by_name

WORLD = 2
GLOBAL_B = 8


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    torch.set_num_threads(1)
    try:
        ocfg = O.OracleConfig(**TINY)
        Loader = get_dataloader_by_name("stroke3-synthetic")
        hps = Loader.parse_hparams(f"max_seq_len={ocfg.max_seq_len}")
        ds = Loader(hps, None, n_classes=ocfg.n_classes, seed=99, rank=rank, world_size=WORLD, vocab=ocfg.vocab_size)
        x, y = next(ds.batch_iterator("train", GLOBAL_B // WORLD))
        # every rank sees the rows [rank*B, (rank+1)*B) of one global batch
        xs = [torch.zeros(GLOBAL_B // WORLD, ocfg.max_seq_len, dtype=torch.int64) for _ in range(WORLD)]
        dist.all_gather(xs, torch.from_numpy(x))
        full = Loader(hps, None, n_classes=ocfg.n_classes, seed=99, rank=0, world_size=1, vocab=ocfg.vocab_size)
        gx, gy = next(full.batch_iterator("train", GLOBAL_B))
        assert np.array_equal(torch.cat(xs).numpy(), gx), "shards are not the slices of the global batch"
        # replicated parameters, local gradients, all-reduce(sum) over a flat arena, 1/world in the optimizer
        p = O.init_params(ocfg, 0, torch.float64)
        _, grads = O.train_step({k: v.clone() for k, v in p.items()}, O.AdamTF(p, ocfg.d_model), torch.from_numpy(x),
                                torch.from_numpy(y), ocfg, dtype=torch.float64)
        # the library's own arena layout (skf_param_table) and bucket ranges (skf_grad_buckets: what skf_train_step
        # all-reduces, bucket by bucket, in backward-completion order) — host code of libskf.so, no GPU involved
        from sketchformer_b200.engine import fill_arena, grad_buckets, make_cfg, param_table
        cfg = make_cfg(GLOBAL_B // WORLD, ocfg.max_seq_len, ocfg.d_model, ocfg.dff, ocfg.num_heads, ocfg.num_layers,
                       ocfg.vocab_size, ocfg.n_classes, ocfg.lowerdim, world_size=WORLD)
        table, total = param_table(cfg)
        arena = fill_arena(torch.zeros(total, dtype=torch.float64), table, ocfg.d_model, grads)
        _, gfull = O.train_step({k: v.clone() for k, v in p.items()}, O.AdamTF(p, ocfg.d_model), torch.from_numpy(gx),
                                torch.from_numpy(gy), ocfg, dtype=torch.float64)
        want = fill_arena(torch.zeros(total, dtype=torch.float64), table, ocfg.d_model, gfull)
        for nb in (1, 2, 3, 4):
            buckets = grad_buckets(cfg, nb)
            assert len(buckets) == nb and buckets[0][1] == total and buckets[-1][0] == 0
            assert all(buckets[i][0] == buckets[i + 1][1] for i in range(nb - 1)), buckets   # contiguous, end -> start
            assert all(lo % 8 == 0 for lo, _ in buckets)                                     # 32-byte aligned ranges
            red = arena.clone()
            for lo, hi in buckets:
                dist.all_reduce(red[lo:hi])
            red /= WORLD
            err = float((red - want).norm() / want.norm())
            assert err < 1e-12, (nb, err)
        # the output layer / decoder come first (backward order), the encoder embedding last
        b4 = grad_buckets(cfg, 4)
        assert b4[0][0] <= table["out.k"][0] < b4[0][1] and b4[-1][0] <= table["enc.emb"][0] < b4[-1][1]
        if rank == 0:
            out.put(("ok", err))
    except Exception as e:      # noqa: BLE001
        out.put(("fail", repr(e)))
        raise
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_average_equals_full_batch():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, port, out)) for r in range(WORLD)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    status, info = out.get(timeout=5)
    assert status == "ok", info


# ---- the plugin's own data-parallel path (sketchformer_b200/model.py) on two gloo ranks ---------------------------------
def _plugin_worker(rank, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    torch.set_num_threads(1)
    try:
        from sketchformer_b200.model import get_model_by_name
        from tests import fake_engine
        fake_engine.install()                      # no GPU here: the engine's arithmetic is the test double's
        Model = get_model_by_name("sketch-transformer-tf2")
        Loader = get_dataloader_by_name("stroke3-synthetic")
        hps = Loader.parse_hparams("max_seq_len=32")
        mh = "num_layers=1,d_model=64,dff=64,num_heads=4,lowerdim=16,dropout_rate=0.0"

        def run(world, r, batch, n_samples):
            ds = Loader(hps, None, n_classes=7, n_samples=n_samples, seed=5, rank=r, world_size=world, vocab=52)
            # dp_buckets=0: the torch.distributed path (all_reduce of the arena after the step, then the optimizer)
            m = Model(Model.parse_hparams(base=f"batch_size={batch},num_epochs=1", specific=mh), ds, None, "dp",
                      device="cpu", world_size=world, rank=r, seed=0, dp_buckets=0, use_cuda_graph=False)
            it = ds.batch_iterator("train", batch)
            for _ in range(3):
                m.train_on_batch(next(it))
            return m

        dp = run(WORLD, rank, GLOBAL_B // WORLD, 64)
        # an epoch is n_samples // (batch_size * world) steps (core/models.py:60 counts one process)
        assert dp.batches_per_epoch == 64 // GLOBAL_B and dp.current_step == 3
        eng = dp._train_engine
        assert [c[2] for c in eng.calls if c[0] == "train_step"] == [False] * 3      # Adam runs AFTER the all-reduce
        assert eng.step_count() == 3
        mine = eng.params.clone()
        both = [torch.zeros_like(mine) for _ in range(WORLD)]
        dist.all_gather(both, mine)
        assert torch.equal(both[0], both[1]), "ranks diverged"
        if rank == 0:
            solo = run(1, 0, GLOBAL_B, 64)._train_engine.params        # one process, the whole global batch
            err = float((mine - solo).norm() / solo.norm())
            assert err < 1e-6, err
            out.put(("ok", err))
    except Exception as e:      # noqa: BLE001
        out.put(("fail", repr(e)))
        raise
    finally:
        dist.destroy_process_group()


def test_plugin_data_parallel_step_on_two_gloo_ranks():
    """Transformer(world_size=2).train_on_batch on two CPU ranks (gloo) with the engine replaced by tests/fake_engine.py:
    every rank trains on its slice of the global batches, the gradient arena is all-reduced by torch.distributed, the
    optimizer runs after it with 1/world, the ranks stay bit-identical and equal to a single process that sees the
    whole global batch."""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_plugin_worker, args=(r, port, out)) for r in range(WORLD)]
    for p in procs:
        p.start()
    status = out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
    assert status[0] == "ok", status

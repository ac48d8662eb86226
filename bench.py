#!/usr/bin/env python
"""bench.py — sketches/s of the Sketchformer hot path on synthetic strokes (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--config cfg2|cfg4|cfg5|fwd|decode] [--scaling weak|strong] [--dp-buckets N]

Default (the headline, BASELINE.json configs[1]): the full train step (fwd + losses + bwd + Adam [+ gradient
all-reduce]) at d_model=256 dff=1024 4-layer 8-head, batch=128 per GPU, max_seq_len=200, V=1004, 345 classes.
Other --config values are the remaining measurement rows of SURVEY.md 8(d):
    cfg4   the same dimensions with use_continuous_data=True (stroke-5 input, Dense(5) head)       [configs[3]]
    cfg5   d_model=512 dff=2048 6-layer 8-head, batch=256, max_seq_len=512                         [configs[4]]
    fwd    forward only, predict_class-style (encoder + bottleneck + classifier), cfg2 dimensions
    decode greedy reconstruction (predict_from_embedding, KV-cached), cfg2 dimensions: sketches/s and tokens/s
N>1 is launched by torchrun (one rank per GPU, NCCL).  --scaling weak: 128 sketches per GPU (the headline);
strong: ONE global batch of 128 split over the ranks.  Prints ONE JSON line (rank 0).  DESIGN.md section 6 explains
every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIMS = {
    "cfg2": dict(num_layers=4, d_model=256, dff=1024, num_heads=8, lowerdim=256, max_seq_len=200, vocab_size=1004,
                 n_classes=345),
    "cfg5": dict(num_layers=6, d_model=512, dff=2048, num_heads=8, lowerdim=256, max_seq_len=512, vocab_size=1004,
                 n_classes=345),
}
DIMS["cfg4"] = DIMS["fwd"] = DIMS["decode"] = DIMS["cfg2"]
BATCHES = {"cfg2": 128, "cfg4": 128, "cfg5": 256, "fwd": 128, "decode": 128}
CFG2, BATCH = DIMS["cfg2"], 128


def workload_name(config, batch):
    c = DIMS[config]
    dims = (f"d_model={c['d_model']} dff={c['dff']} {c['num_layers']}-layer {c['num_heads']}-head "
            f"lowerdim={c['lowerdim']}, batch={batch}/GPU, max_seq_len={c['max_seq_len']}")
    return {
        "cfg2": f"cfg2 train step: {dims}, dict-token V=1004, 345 classes, dropout=0, fwd+losses+bwd+Adam",
        "cfg4": f"cfg4 train step (use_continuous_data=True: stroke-5 input, Dense(5) head, MSE+CE loss): {dims}, "
                "345 classes, dropout=0, fwd+losses+bwd+Adam",
        "cfg5": f"cfg5 train step: {dims}, dict-token V=1004, 345 classes, dropout=0, fwd+losses+bwd+Adam",
        "fwd": f"cfg2 forward only (predict_class: encoder + bottleneck + classifier argmax): {dims}, dict-token V=1004",
        "decode": f"cfg2 greedy decode (predict_from_embedding, <= max_seq_len KV-cached steps, blind decoder): {dims}",
    }[config]


def core_flops_per_sketch(T=200, d=256, f=1024, L=4, train=True, encoder_only=False):
    """algorithmic attention + FFN FLOPs per sketch (2mnk, causal self-attention at 1/2)."""
    Te, Td = T, T - 1
    enc = 8 * Te * d * d + 4 * Te * Te * d + 4 * Te * d * f
    dec = (8 * Td * d * d + 4 * (Td * (Td + 1) // 2) * d + 4 * Td * d * d + 4 * Te * d * d + 4 * Td * Te * d
           + 4 * Td * d * f)
    fwd = L * (enc if encoder_only else enc + dec)
    return 3 * fwd if train else fwd


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, False, []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None, "reasons": reasons,
                "samples": len(self.rows)}


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle (fp32 torch-CPU restatement of the TF-2.1 graph) on the host cores
# ---------------------------------------------------------------------------------------------------
def cpu_reference_steps(config, steps, warmup, budget_s, batch=None):
    """Times `steps` passes of the oracle over a bounded sample of the workload (batch <= the config's batch, chosen
    from a probe so that the whole run fits `budget_s`); sketches/s are linear in the batch for every pass here."""
    import torch
    from oracle import sketchformer_oracle as O
    from sketchformer_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cont = config == "cfg4"
    ocfg = O.OracleConfig(**DIMS[config], use_continuous_data=cont)
    p = O.init_params(ocfg, 0, torch.float32)
    opt = O.AdamTF(p, ocfg.d_model)
    full = BATCHES[config]

    def data(b):
        if cont:
            x, y = synth.continuous_batch(b, ocfg.max_seq_len, ocfg.n_classes, 1234)
        else:
            x, y = synth.token_batch(b, ocfg.max_seq_len, ocfg.n_classes, 1234)
        return torch.from_numpy(x), torch.from_numpy(y)

    def run(b, n):
        xt, yt = data(b)
        emb = None
        if config == "decode":
            with torch.no_grad():
                emb = O.predict_class(p, xt, ocfg, torch.float32)["embedding"]
        t0 = time.perf_counter()
        tokens = 0
        for _ in range(n):
            if config == "fwd":
                with torch.no_grad():
                    O.predict_class(p, xt, ocfg, torch.float32)
            elif config == "decode":
                # models/sketchformer.py:274-301: the reference re-runs the decoder on the growing prefix every step
                with torch.no_grad():
                    out = O.predict_from_embedding(p, emb, ocfg, None, 1002, 1003, torch.float32)
                tokens += int(out["recon"].shape[0] * (out["recon"].shape[1] - 1))
            else:
                O.train_step(p, opt, xt, yt, ocfg, dtype=torch.float32)
        return time.perf_counter() - t0, tokens

    probe = 2 if config == "cfg5" else 4
    if batch is None and config == "decode":
        batch = 4       # 200 sequential full-prefix decoder passes per call (2 and 16 sketches measured the same rate)
    if batch is None:
        run(probe, 1)                           # page in / warm the thread pool
        t, _ = run(probe, 1)
        per_sketch = t / probe
        batch = int(max(probe, min(full, budget_s / max(1, steps + warmup) / per_sketch)))
    if warmup:
        run(batch, warmup)
    dt, tokens = run(batch, steps)
    return {"value": batch * steps / dt, "batch": batch, "seconds": dt, "cores": cores, "tokens_per_s": tokens / dt}


def bench_config(cfg, per_gpu, world):
    """The `config` object of the JSON line — identical in both arms (the driver compares them): what is computed, on
    how many ranks, and how the timed region keeps L2 from serving repeated inputs."""
    return {"workload": workload_name(cfg, per_gpu), "parallelism": f"dp{world}", "global_batch": per_gpu * world,
            "l2": "per-step working set (activations saved for backward, GBs) >> 126 MB L2; 8 distinct batches"}


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = args.config
    heavy = cfg in ("cfg5", "decode")
    steps, warmup = (min(args.steps, 2), min(args.warmup, 1)) if heavy else (args.steps, args.warmup)
    r = cpu_reference_steps(cfg, steps, warmup, budget_s=150.0)
    full = BATCHES[cfg]
    sample = (f"{steps} passes over {r['batch']} of the {full} sketches of the {cfg} batch "
              f"(linear in batch), "
              f"fp32 oracle, torch-CPU {r['cores']} threads")
    line = {"impl": "reference", "metric": "sketches/sec", "value": r["value"], "unit": "sketches/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * r["seconds"] / steps * (full / r["batch"]), "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(cfg, full if args.scaling != "strong" else full // max(1, args.gpus), max(1, args.gpus)),
            "note": "reference TF-2.1 cannot run in this image; this is the fp32 torch-CPU op-for-op restatement "
                    "(oracle/) on the host cores of rank 0",
            "cpu_baseline": {"value": r["value"], "unit": "sketches/s", "cores": r["cores"], "kind": "port",
                             "sample": sample},
            "e2e": {"value": r["value"], "unit": "sketches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if cfg == "decode":
        line["tokens_per_s"] = r["tokens_per_s"]
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# own arm
# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default="cfg2", choices=sorted(BATCHES))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--dp-buckets", type=int, default=-1,
                    help="gradient buckets all-reduced inside the step (0: one torch all_reduce after the step; "
                         "-1: the plugin's default, in-step on 2 ranks and the torch path on more)")
    ap.add_argument("--precision", type=int, default=-1,
                    help="tensor-core GEMM passes: 0 TF32 everywhere, 1 3xTF32 forward, 2 3xTF32 everywhere (default: engine default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-loader", action="store_true", help="skip the stroke3-distributed loader measurement")
    ap.add_argument("--set", action="append", default=[], metavar="NAME=INT",
                    help="library switch for A/B runs (skf_set_option), e.g. --set wgrad_overlap=0")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return reference_main(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from sketchformer_b200.data import get_dataloader_by_name
    from sketchformer_b200.model import get_model_by_name

    for kv in args.set:
        from sketchformer_b200 import _lib
        k, v = kv.split("=")
        _lib.check(_lib.load().skf_set_option(k.encode(), int(v)), "skf_set_option")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL's version / warning lines go to stderr: stdout carries the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    assert world == args.gpus or world == 1 and args.gpus == 1, "launch with torchrun for --gpus > 1"

    cfg = args.config
    c = DIMS[cfg]
    train = cfg in ("cfg2", "cfg4", "cfg5")
    full = BATCHES[cfg]
    if args.scaling == "strong":
        assert full % world == 0, "strong scaling splits the global batch evenly"
        B = full // world
    else:
        B = full
    # the reference-facing plugin API: registry names and hparam strings exactly as train.py passes them
    Model = get_model_by_name("sketch-transformer-tf2")
    Loader = get_dataloader_by_name("stroke3-synthetic")
    data_hps = Loader.parse_hparams(f"max_seq_len={c['max_seq_len']},use_continuous_data={cfg == 'cfg4'}")
    model_hps = Model.parse_hparams(base=f"batch_size={B}",
                                    specific=f"num_layers={c['num_layers']},d_model={c['d_model']},dff={c['dff']},"
                                             f"num_heads={c['num_heads']},lowerdim={c['lowerdim']},dropout_rate=0.0")
    if args.scaling == "strong":      # every rank draws the same global batch and keeps its slice
        dataset = Loader(data_hps, None, n_classes=c["n_classes"], seed=1234, rank=rank, world_size=world)
    else:
        dataset = Loader(data_hps, None, n_classes=c["n_classes"], seed=1234 + 1000 * rank)
    model = Model(model_hps, dataset, None, "bench", device=f"cuda:{local}", world_size=world, seed=0, rank=rank,
                  gemm_precision=None if args.precision < 0 else args.precision,
                  dp_buckets=None if args.dp_buckets < 0 else args.dp_buckets)
    args.dp_buckets = model._dp_buckets
    eng = model._train_engine           # identical random-init weights on every rank (Keras initialisers, seed 0)

    NB = 8   # distinct synthetic batches per rank: resident in HBM (value) / host numpy arrays (e2e)
    it = dataset.batch_iterator("train", B)
    host = [next(it) for _ in range(NB)]
    dev_x = [torch.from_numpy(x).cuda() for x, _ in host]
    dev_y = [torch.from_numpy(y).cuda() for _, y in host]
    last = {}
    tokens = [0]
    dp_check = dp_gradient_check(model, world, rank, full, c, cfg) if (world > 1 and train) else None

    if cfg == "decode":
        emb = [np.ascontiguousarray(model.predict_class(x)["embedding"]) for x, _ in host]
        dev_emb = [torch.from_numpy(e).cuda() for e in emb]
        sos, eos = dataset.tokenizer.SOS, dataset.tokenizer.EOS

    def step(i, e2e=False):
        j = i % NB
        if train:
            if e2e:       # public API, host buffers in, metrics dict (floats on the host) out
                last.update(model.train_on_batch(host[j]))
            else:         # inputs already resident in HBM, no device->host read
                model.train_on_batch((dev_x[j], dev_y[j]), sync_metrics=False)
        elif cfg == "fwd":
            if e2e:       # host tokens in, class ids + embeddings (host arrays) out
                last.update(model.predict_class(host[j][0]))
            else:
                eng.encode(dev_x[j])
        else:
            if e2e:       # host embeddings in, token sequences (host array) out
                out = model.predict_from_embedding(emb[j])
                tokens[0] += out["recon"].shape[0] * (out["recon"].shape[1] - 1)
            else:
                out = eng.decode(dev_emb[j], None, sos, eos)
                tokens[0] += out.shape[0] * (out.shape[1] - 1)

    host_ms = [0.0]

    def timed(n, e2e=False):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        tokens[0] = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        for i in range(n):
            step(i, e2e)
        e1.record()
        host_ms[0] = (time.perf_counter() - t0) * 1e3      # host submission time (diagnostic only)
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    steps = args.steps if cfg != "decode" else min(args.steps, 10)
    lw = eng.launches()
    step(0)                               # eager: its launch count is what a (graph-replayed) call executes
    launches_per_call = eng.launches() - lw
    for i in range(1, args.warmup):       # >= 3: one eager step, then the CUDA graph is captured and replayed
        step(i)
    l0 = eng.launches()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(steps)
    tok_dev = tokens[0]
    host_submit_ms = host_ms[0] / steps
    eager_launches = eng.launches() - l0
    for i in range(2):
        step(i, True)
    ms_e2e = timed(steps, e2e=True)
    tok_e2e = tokens[0]
    sampler.stop_flag = True
    if train:
        launches = (getattr(eng, "graph_launches", 0) + (0 if world == 1 or args.dp_buckets else 2)) * steps
    else:
        launches = max(eager_launches, launches_per_call * steps)

    value = world * B * steps / (ms / 1e3)
    e2e_value = world * B * steps / (ms_e2e / 1e3)
    peaks, peak_src = load_peaks()
    precision = int(eng.cfg.gemm_precision)

    # roofline of the dominant kernel (the tcgen05 pair GEMM): CUDA events around launches of this config's FFN-1
    # shape through the C ABI, on this stream, after the timed region
    roof = gemm_roofline(B, c, peaks, peak_src, 5 if precision >= 1 else 0)
    core = core_flops_per_sketch(c["max_seq_len"], c["d_model"], c["dff"], c["num_layers"], train=train,
                                 encoder_only=cfg == "fwd")
    if cfg != "decode":
        roof["step_core_tflops"] = value / world * core / 1e12
        roof["step_frac_of_bf16_peak"] = roof["step_core_tflops"] / peaks["bf16_tflops_sustained"]
    else:
        roof = decode_roofline(eng, c, B, ms / steps, tok_dev / steps, peaks, peak_src)

    if rank == 0:
        mode = {0: "tf32 single pass everywhere", 1: "3xTF32 forward, TF32 backward",
                2: "error-compensated 3-term products (fp32-grade) forward and backward: bf16x3 in the pair GEMM "
                   "(bf16(A)bf16(B) + bf16(A)bf16(Bl) + bf16(Al)bf16(B), fp32 accumulate) and in attention",
                3: "3xTF32 forward and dgrad, TF32 weight gradients"}.get(precision, str(precision))
        if cfg == "cfg4":
            h2d = B * c["max_seq_len"] * 5 * 4 + B * 8
        elif cfg == "decode":
            h2d = B * c["d_model"] * 4
        else:
            h2d = B * c["max_seq_len"] * 8 + (B * 8 if train else 0)
        d2h = 128 if train else (B * 4 + B * c["d_model"] * 4 if cfg == "fwd" else B * (c["max_seq_len"] + 1) * 4 + B * 4)
        line = {"metric": "sketches/sec", "value": value, "unit": "sketches/s", "n_gpus": world,
                "steps": steps, "warmup": args.warmup, "ms_per_step": ms / steps,
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": "f32 (tensor-core split products: bf16x3 GEMMs and attention; fp32 accumulate and storage)",
                "data": "synthetic",
                "config": bench_config(cfg, B, world),
                # how this arm ran it (not part of the workload definition the two arms share)
                "run": {"precision": mode, "cuda_graph": bool(getattr(eng, "_graph", None)),
                        "dp_exchange": ("none" if world == 1 else
                                        f"{args.dp_buckets} NCCL buckets inside the step (graph-captured, overlapped with backward)"
                                        if args.dp_buckets else "one torch.distributed all_reduce after the step"),
                        "api": "sketchformer_b200.model.Transformer." +
                               ("train_on_batch (reference models/sketchformer.py:351)" if train else
                                "predict_class (reference models/sketchformer.py:223)" if cfg == "fwd" else
                                "predict_from_embedding (reference models/sketchformer.py:255)")},
                "e2e": {"value": e2e_value, "unit": "sketches/s", "ms_per_step": ms_e2e / steps,
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": launches, "host_submit_ms_per_step": host_submit_ms,
                "clocks": sampler.summary(), "roofline": roof}
        if train:
            line["e2e"]["last_total_loss"] = last.get("total_loss")
        if cfg == "decode":
            line["tokens_per_s"] = world * tok_dev / (ms / 1e3)
            line["e2e"]["tokens_per_s"] = world * tok_e2e / (ms_e2e / 1e3)
            line["decode_steps_per_call"] = tok_dev / steps / B
        if dp_check is not None:
            line["dp_grad_rel_err"] = dp_check
        if cfg == "cfg2" and world == 1 and not args.no_loader:
            line["e2e_loader"] = loader_bench(model, c, steps)
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_steps(cfg, 2 if cfg != "decode" else 1, 1 if cfg not in ("cfg5", "decode") else 0,
                                    budget_s=25.0)
            line["cpu_baseline"] = {"value": r["value"], "unit": "sketches/s", "cores": r["cores"], "kind": "port",
                                    "sample": f"oracle fp32 on {r['batch']} of the {full} sketches (linear in batch), "
                                              f"torch-CPU {r['cores']} threads"}
            if cfg == "decode":
                line["cpu_baseline"]["tokens_per_s"] = r["tokens_per_s"]
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def loader_bench(model, c, steps):
    """the real `stroke3-distributed` iterator (chunked .npz files, prefetch thread, whole-chunk
    tokenisation on the device, batches cut by fancy indexing) over a synthetic chunk directory in the reference's
    on-disk format, alone and feeding train_on_batch.  Reference: dataloaders/distributed_stroke3.py:77-125,
    core/data.py:79-172 (a per-sketch Python loop and per-sample appends)."""
    import shutil
    import tempfile
    import numpy as np
    import torch
    from sketchformer_b200.data import get_dataloader_by_name
    from sketchformer_b200.tokenizer import Tokenizer
    root = tempfile.mkdtemp(prefix="bench_loader_", dir=ROOT)
    try:
        rng = np.random.default_rng(7)
        per_chunk, n_chunks = 8192, 4

        def chunk(n):
            lens = rng.integers(20, 121, size=n)
            flat = np.zeros((int(lens.sum()), 3), np.int16)
            flat[:, :2] = rng.integers(-40, 41, size=(len(flat), 2))
            flat[:, 2] = rng.random(len(flat)) < 0.08
            ends = np.cumsum(lens)
            flat[ends - 1, 2] = 1
            x = np.empty(n, dtype=object)
            for i, part in enumerate(np.split(flat, ends[:-1])):
                x[i] = part
            return x, rng.integers(0, c["n_classes"], size=n)
        for i in range(n_chunks):
            x, y = chunk(per_chunk)
            np.savez(os.path.join(root, f"train_{i:03d}.npz"), x=x, y=y)
        for name in ("valid", "test"):
            x, y = chunk(256)
            np.savez(os.path.join(root, f"{name}.npz"), x=x, y=y)
        np.savez(os.path.join(root, "meta.npz"), n_classes=c["n_classes"], n_samples_train=per_chunk * n_chunks,
                 n_samples_valid=256, n_samples_test=256, class_names=np.array([f"c{i}" for i in range(c["n_classes"])]),
                 std=1.0, mean=0.0)
        Loader = get_dataloader_by_name("stroke3-distributed")
        hps = Loader.parse_hparams(f"max_seq_len={c['max_seq_len']}")
        cents = np.random.default_rng(0).uniform(-1, 1, size=(1000, 2)).astype(np.float32)
        ds = Loader(hps, root, seed=5, tokenizer=Tokenizer(centroids=cents))
        out = {"unit": "sketches/s", "chunks": n_chunks, "sketches_per_chunk": per_chunk,
               "device_tokenizer": bool(ds._device_path())}
        # one chunk through preprocess(): device path, and the host restatement on a sample of it
        raw = np.load(os.path.join(root, "train_000.npz"), allow_pickle=True)["x"]
        ds.preprocess(raw[:256])
        t0 = time.perf_counter(); ds.preprocess(raw); dt = time.perf_counter() - t0
        out["preprocess_sketches_per_s"] = len(raw) / dt
        dev, ds.device = ds.device, None
        t0 = time.perf_counter(); ds.preprocess(raw[:512]); dt = time.perf_counter() - t0
        out["preprocess_host_loop_sketches_per_s"] = 512 / dt
        ds.device = dev
        # the iterator alone (2.3 passes over the chunks: swaps and the wrap-around included)
        it = ds.batch_iterator("train", BATCH)
        for _ in range(3):
            next(it)
        nb = 600
        t0 = time.perf_counter()
        for _ in range(nb):
            next(it)
        out["loader_sketches_per_s"] = nb * BATCH / (time.perf_counter() - t0)
        # and feeding the train step through the plugin API
        for _ in range(3):
            model.train_on_batch(next(it))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            model.train_on_batch(next(it), sync_metrics=False)
        torch.cuda.synchronize()
        out["value"] = steps * BATCH / (time.perf_counter() - t0)
        return out
    finally:
        shutil.rmtree(root, ignore_errors=True)


def dp_gradient_check(model, world, rank, full, c, cfg):
    """the ranks take the slices of ONE global batch of `full` sketches; the all-reduced gradient
    arena (sum of per-rank mean gradients, times 1/world) must equal rank 0's gradient of the whole batch.
    Returns the relative L2 error on rank 0 (None elsewhere).  One untimed step before the benchmark."""
    import torch
    import torch.distributed as dist
    from sketchformer_b200 import synth
    if full % world:
        return None
    per = full // world
    gen = synth.continuous_batch if cfg == "cfg4" else synth.token_batch
    x, y = gen(full, c["max_seq_len"], c["n_classes"], 4242)
    xs, ys = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    part = model._engine(per, training=True)
    part.set_comm(None)
    part.train_step(xs[rank * per:(rank + 1) * per].contiguous(), ys[rank * per:(rank + 1) * per].contiguous(),
                    apply_adam=False)
    comm = model._comm
    if comm is None:
        dist.all_reduce(part.grads)
    else:
        part.allreduce_grads(comm)
    torch.cuda.synchronize()
    reduced = (part.grads / world).clone()
    err = None
    if rank == 0:
        whole = model._engine(full, training=True)
        whole.set_comm(None)
        whole.train_step(xs, ys, apply_adam=False)
        torch.cuda.synchronize()
        err = float((reduced.double() - whole.grads.double()).norm() / whole.grads.double().norm())
        if model._dp_buckets:
            whole.set_comm(model._comm, model._dp_buckets)
    if model._dp_buckets:
        part.set_comm(model._comm, model._dp_buckets)
    dist.barrier()
    return err


def gemm_traffic_from_capture():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE gemm3_k launch of the cfg2 FFN-1 shape, parsed from the
    newest committed `ncu --set full` summary under profiles/ (written by tools/ncu_summary.py)."""
    import glob
    import re
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*ncu_full_gemm3_pair_ffn1.txt"))):
        rd = wr = None
        for ln in open(path):
            m = re.search(r"dram__bytes_(read|write)\.sum\s+([0-9.,]+)\s+(\w+)", ln)
            if m:
                v = float(m.group(2).replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(m.group(3), 1)
                if m.group(1) == "read" and rd is None:
                    rd = v
                if m.group(1) == "write" and wr is None:
                    wr = v
        if rd is not None and wr is not None:
            best = (rd + wr, os.path.basename(path))
    return best


def gemm_roofline(B, c, peaks, peak_src, backend):
    """achieved TFLOP/s of the dominant kernel on its own stream (CUDA events), FFN-1 shape of the config:
    [B*T, d] x [d, dff] + bias + ReLU (cfg2: 2*25600*256*1024 = 13.42 algorithmic GFLOP per launch).
    backend 5 = bf16x3 on CTA pairs (gemm3_k, three tensor-core passes per algorithmic FLOP, what the plan runs),
    0 = single-pass TF32."""
    import torch
    from sketchformer_b200.ops import gemm, split_bf16
    M, K, N = B * c["max_seq_len"], c["d_model"], c["dff"]
    X = torch.randn(M, K, device="cuda")
    W = torch.randn(K, N, device="cuda")
    bias = torch.zeros(N, device="cuda")
    nbuf = max(2, min(4, int(2e9 // (M * N * 4))))
    Ys = [torch.empty(M, N, device="cuda") for _ in range(nbuf)]    # outputs cycling over > L2
    # as in the plan's forward pass: the weight's bf16 hi / lo images exist (one split pass per step) and the pair
    # kernel TMA-loads its B tiles from them
    img = dict(zip(("b_hi16", "b_lo16"), split_bf16(W))) if backend == 5 else {}
    for i in range(3):
        gemm(X, K, 1, W, N, 1, Ys[i % nbuf], N, M, N, K, backend, bias=bias, act=1, **img)
    torch.cuda.synchronize()
    n = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        gemm(X, K, 1, W, N, 1, Ys[i % nbuf], N, M, N, K, backend, bias=bias, act=1, **img)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    flops = 2.0 * M * N * K
    achieved = flops / (us * 1e-6) / 1e12
    # The pair kernel (bf16x3) runs three bf16 MMA passes per algorithmic FLOP: quoted against the measured bf16
    # burst peak; the single-CTA kernels run TF32 MMAs (half the bf16 rate).
    peak = peaks["bf16_tflops"] if backend == 5 else peaks["bf16_tflops"] / 2
    passes = 3 if backend == 5 else 1
    kname = {5: "gemm3_k<256,false,true,2,true> (cta_group::2 pair; bf16x3: three kind::f16 MMA passes; A split into "
                "bf16 hi/lo tiles in shared memory, B TMA-loaded from the pre-split weight images)",
             0: "gemm2_k<256,false,true,false> (TF32)"}[backend]
    traffic = gemm_traffic_from_capture() if (M, K, N) == (25600, 256, 1024) and backend == 5 else None
    return {"bound": "tensor", "kernel": kname + f" FFN-1: {M}x{K}x{N}, bias+ReLU epilogue",
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "traffic": traffic[0] if traffic else None, "traffic_source": traffic[1] if traffic else None,
            "us_per_launch": us, "flops_per_launch": flops, "mma_passes": passes,
            "tensor_pipe_tflops": achieved * passes, "tensor_pipe_frac": achieved * passes / peak,
            "peak_source": peak_src + ("; bf16 burst figure" if backend == 5 else "; TF32 = 1/2 of the bf16 burst figure"),
            "hbm_algorithmic_bytes": 4.0 * (M * K + K * N + M * N)}


def decode_roofline(eng, c, B, ms_per_call, tokens_per_call, peaks, peak_src):
    """Greedy decode is latency bound (one token per sequential step, ~50 small launches each); the bound that would
    apply to a perfect implementation is HBM: every step re-reads the decoder's weights and the key / value caches."""
    d, f, L, T, V = c["d_model"], c["dff"], c["num_layers"], c["max_seq_len"], c["vocab_size"]
    w = L * (4 * d * d + 2 * d * d + 2 * d * f) + d * V + V * d          # decoder + output layer + embedding rows
    steps = max(1.0, tokens_per_call / B)
    kv = L * B * (steps / 2) * 2 * d * 4 + L * B * T * 2 * d * 4           # self cache (half full on average) + cross K/V
    byts = steps * (w * (4 + 4) + kv)                                      # fp32 weights + their bf16 hi/lo images
    achieved = byts / (ms_per_call * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": "greedy decode step (decode_attn_k + per-token GEMMs, launch-latency bound)",
            "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
            "traffic": None, "bytes_per_call": byts, "steps_per_call": steps, "peak_source": peak_src}


if __name__ == "__main__":
    main()

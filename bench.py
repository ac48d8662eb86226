#!/usr/bin/env python
"""bench.py — sketches/s of the Sketchformer train step (fwd + losses + bwd + Adam [+ grad all-reduce])
on synthetic dict-tokenised strokes, BASELINE.json configs[1]: d_model=256 dff=1024 4-layer 8-head,
batch=128 per GPU, max_seq_len=200, V=1004, 345 classes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N>1 is launched by torchrun (one rank per GPU, NCCL); weak scaling (128 sketches per GPU).
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
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

CFG2 = dict(num_layers=4, d_model=256, dff=1024, num_heads=8, lowerdim=256, max_seq_len=200, vocab_size=1004,
            n_classes=345)
BATCH = 128
WORKLOAD = ("cfg2 train step: d_model=256 dff=1024 4-layer 8-head lowerdim=256, batch=128/GPU, max_seq_len=200, "
            "dict-token V=1004, 345 classes, dropout=0, fwd+losses+bwd+Adam")
# algorithmic core (attention+FFN) FLOPs, train = 3 x forward, per sketch
F_TRAIN_GFLOP = 10.011


def core_flops_per_sketch(T=200, d=256, f=1024, L=4):
    Te, Td = T, T - 1
    enc = 8 * Te * d * d + 4 * Te * Te * d + 4 * Te * d * f
    dec = (8 * Td * d * d + 4 * (Td * (Td + 1) // 2) * d + 4 * Td * d * d + 4 * Te * d * d + 4 * Td * Te * d
           + 4 * Td * d * f)
    return 3 * L * (enc + dec)


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
# reference arm / cpu_baseline: the oracle (fp32 torch-CPU restatement of the TF-2.1 graph)
# ---------------------------------------------------------------------------------------------------
def cpu_reference_steps(steps, warmup, budget_s, batch=None):
    """Times `steps` oracle train steps on a bounded sample (batch <= 128 sketches of cfg2)."""
    import torch
    from oracle import sketchformer_oracle as O
    from sketchformer_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ocfg = O.OracleConfig(**CFG2)
    p = O.init_params(ocfg, 0, torch.float32)
    opt = O.AdamTF(p, ocfg.d_model)

    def run(b, n):
        x, y = synth.token_batch(b, ocfg.max_seq_len, ocfg.n_classes, 1234)
        xt, yt = torch.from_numpy(x), torch.from_numpy(y)
        t0 = time.perf_counter()
        for _ in range(n):
            O.train_step(p, opt, xt, yt, ocfg, dtype=torch.float32)
        return time.perf_counter() - t0

    if batch is None:
        run(4, 1)                           # page in / warm the thread pool
        t4 = run(4, 1)
        per_sketch = t4 / 4
        batch = int(max(4, min(BATCH, budget_s / max(1, steps + warmup) / per_sketch)))
    run(batch, warmup)
    dt = run(batch, steps)
    return {"value": batch * steps / dt, "batch": batch, "seconds": dt, "cores": cores}


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_steps(args.steps, args.warmup, budget_s=150.0)
    sample = f"{args.steps} train steps on {r['batch']} of the {BATCH} sketches of the cfg2 batch (linear in batch)"
    line = {"impl": "reference", "metric": "sketches/sec", "value": r["value"], "unit": "sketches/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * r["seconds"] / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "reference TF-2.1 cannot run in this image; this is the "
                       "fp32 torch-CPU op-for-op restatement (oracle/) on the host cores"},
            "cpu_baseline": {"value": r["value"], "unit": "sketches/s", "cores": r["cores"], "kind": "port",
                             "sample": sample},
            "e2e": {"value": r["value"], "unit": "sketches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
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
    ap.add_argument("--backend", type=int, default=0, help="0 tcgen05 TF32 GEMMs, 1 SIMT fp32 (debug)")
    ap.add_argument("--precision", type=int, default=-1,
                    help="tensor-core GEMM passes: 0 TF32 everywhere, 1 3xTF32 forward, 2 3xTF32 everywhere (default: engine default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--set", action="append", default=[], metavar="NAME=INT",
                    help="library switch for A/B runs (skf_set_option), e.g. --set wgrad_overlap=0")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return reference_main(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from sketchformer_b200 import synth
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

    # the reference-facing plugin API: registry names and hparam strings exactly as train.py passes them
    c = CFG2
    Model = get_model_by_name("sketch-transformer-tf2")
    Loader = get_dataloader_by_name("stroke3-synthetic")
    data_hps = Loader.parse_hparams(f"max_seq_len={c['max_seq_len']}")
    model_hps = Model.parse_hparams(base=f"batch_size={BATCH}",
                                    specific=f"num_layers={c['num_layers']},d_model={c['d_model']},dff={c['dff']},"
                                             f"num_heads={c['num_heads']},lowerdim={c['lowerdim']},dropout_rate=0.0")
    dataset = Loader(data_hps, None, n_classes=c["n_classes"], seed=1234 + 1000 * rank)
    model = Model(model_hps, dataset, None, "bench", device=f"cuda:{local}", world_size=world, seed=0, rank=rank,
                  gemm_precision=None if args.precision < 0 else args.precision)
    eng = model._train_engine           # identical random-init weights on every rank (Keras initialisers, seed 0)

    NB = 8   # distinct synthetic batches per rank: resident in HBM (value) / host numpy arrays (e2e)
    it = dataset.batch_iterator("train", BATCH)
    host = [next(it) for _ in range(NB)]
    dev_x = [torch.from_numpy(x).cuda() for x, _ in host]
    dev_y = [torch.from_numpy(y).cuda() for _, y in host]
    last = {}

    def step(i, e2e=False):
        if e2e:       # public API, host buffers in, metrics dict (floats on the host) out
            last.update(model.train_on_batch(host[i % NB]))
        else:         # inputs already resident in HBM, no device->host read
            model.train_on_batch((dev_x[i % NB], dev_y[i % NB]), sync_metrics=False)

    host_ms = [0.0]

    def timed(n, e2e=False):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
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

    for i in range(args.warmup):          # >= 3: two eager steps, then the CUDA graph is captured and replayed
        step(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(args.steps)
    host_submit_ms = host_ms[0] / args.steps
    for i in range(2):
        step(i, True)
    ms_e2e = timed(args.steps, e2e=True)
    sampler.stop_flag = True
    launches_per_step = getattr(eng, "graph_launches", 0) + (0 if world == 1 else 2)

    value = world * BATCH * args.steps / (ms / 1e3)
    e2e_value = world * BATCH * args.steps / (ms_e2e / 1e3)
    peaks, peak_src = load_peaks()
    precision = int(eng.cfg.gemm_precision)

    # roofline of the dominant kernel (the persistent tcgen05 GEMM): CUDA events around launches of the FFN-1
    # shape through the C ABI, on this stream, after the timed region
    roof = gemm_roofline(eng, c, peaks, peak_src, 5 if precision >= 1 else 0)
    roof["step_core_tflops"] = value / world * core_flops_per_sketch() / 1e12
    roof["step_frac_of_tf32_peak"] = roof["step_core_tflops"] / (peaks["bf16_tflops_sustained"] / 2)
    roof["step_frac_of_bf16_peak"] = roof["step_core_tflops"] / peaks["bf16_tflops_sustained"]

    if rank == 0:
        mode = {0: "tf32 single pass everywhere", 1: "3xTF32 forward, TF32 backward",
                2: "error-compensated 3-term products (fp32-grade) forward and backward: bf16x3 in the pair GEMM "
                   "(bf16(A)bf16(B) + bf16(A)bf16(Bl) + bf16(Al)bf16(B), fp32 accumulate), 3xTF32 in attention",
                3: "3xTF32 forward and dgrad, TF32 weight gradients"}.get(precision, str(precision))
        line = {"metric": "sketches/sec", "value": value, "unit": "sketches/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (tensor-core split products: bf16x3 GEMMs, 3xTF32 attention; fp32 accumulate and storage)",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "parallelism": f"dp{world}", "global_batch": BATCH * world,
                           "l2": "per-step working set 3.5 GB (activations saved for backward) >> 126 MB L2; 8 distinct batches",
                           "precision": mode, "cuda_graph": bool(getattr(eng, "_graph", None)),
                           "api": "sketchformer_b200.model.Transformer.train_on_batch (reference models/sketchformer.py:351)"},
                "e2e": {"value": e2e_value, "unit": "sketches/s", "ms_per_step": ms_e2e / args.steps,
                        "h2d_bytes_per_step": BATCH * c["max_seq_len"] * 8 + BATCH * 8, "d2h_bytes_per_step": 128,
                        "last_total_loss": last.get("total_loss")},
                "gpu_launches": launches_per_step * args.steps, "host_submit_ms_per_step": host_submit_ms,
                "clocks": sampler.summary(), "roofline": roof}
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_steps(2, 1, budget_s=25.0)
            line["cpu_baseline"] = {"value": r["value"], "unit": "sketches/s", "cores": r["cores"], "kind": "port",
                                    "sample": f"2 oracle fp32 train steps on {r['batch']} of the {BATCH} sketches "
                                              f"(linear in batch), torch-CPU {r['cores']} threads"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum of ONE gemm3_k launch of this shape from the committed
# `ncu --set full` capture (profiles/r1_s7_ncu_full_gemm3_pair_ffn1.txt): 27.3 MB read + 45.6 MB written (part of the
# 105 MB output is still in the 126 MB L2 when the kernel ends); algorithmic bytes are 132 MB.
NCU_TRAFFIC_BYTES = {5: 73.0e6, 2: 72.2e6}


def gemm_roofline(eng, c, peaks, peak_src, backend):
    """achieved TFLOP/s of the dominant kernel on its own stream (CUDA events), FFN-1 shape of cfg2:
    [B*T, d] x [d, dff] + bias + ReLU = 2*25600*256*1024 = 13.42 algorithmic GFLOP per launch.
    backend 5 = 3xTF32 on CTA pairs (gemm3_k, three tensor-core passes per algorithmic FLOP, what the plan
    runs), 2 = the single-CTA 3xTF32 kernel, 0 = single-pass TF32."""
    import torch
    from sketchformer_b200.ops import gemm, split_bf16
    M, K, N = BATCH * c["max_seq_len"], c["d_model"], c["dff"]
    X = torch.randn(M, K, device="cuda")
    W = torch.randn(K, N, device="cuda")
    bias = torch.zeros(N, device="cuda")
    Ys = [torch.empty(M, N, device="cuda") for _ in range(4)]    # 4 x 105 MB outputs: > L2
    # as in the plan's forward pass: the weight's bf16 hi / lo images exist (one split pass per step) and the pair
    # kernel TMA-loads its B tiles from them
    img = dict(zip(("b_hi16", "b_lo16"), split_bf16(W))) if backend == 5 else {}
    for i in range(3):
        gemm(X, K, 1, W, N, 1, Ys[i % 4], N, M, N, K, backend, bias=bias, act=1, **img)
    torch.cuda.synchronize()
    n = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        gemm(X, K, 1, W, N, 1, Ys[i % 4], N, M, N, K, backend, bias=bias, act=1, **img)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    flops = 2.0 * M * N * K
    achieved = flops / (us * 1e-6) / 1e12
    # The pair kernel (bf16x3) runs three bf16 MMA passes per algorithmic FLOP: quoted against the measured bf16
    # burst peak; the single-CTA kernels run TF32 MMAs (half the bf16 rate), three passes in 3xTF32 mode.
    peak = peaks["bf16_tflops"] if backend == 5 else peaks["bf16_tflops"] / 2
    passes = {5: 3, 2: 3}.get(backend, 1)
    kname = {5: "gemm3_k<256,false,true,2,true> (cta_group::2 pair; bf16x3: three kind::f16 MMA passes; A split into "
                "bf16 hi/lo tiles in shared memory, B TMA-loaded from the pre-split weight images)",
             2: "gemm2_k<256,false,true,true> (3xTF32)", 0: "gemm2_k<256,false,true,false> (TF32)"}[backend]
    return {"bound": "tensor", "kernel": kname + " FFN-1: 25600x256x1024, bias+ReLU epilogue",
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "traffic": NCU_TRAFFIC_BYTES.get(backend),
            "us_per_launch": us, "flops_per_launch": flops, "mma_passes": passes,
            "tensor_pipe_tflops": achieved * passes, "tensor_pipe_frac": achieved * passes / peak,
            "peak_source": peak_src + ("; bf16 burst figure" if backend == 5 else "; TF32 = 1/2 of the bf16 burst figure"),
            "hbm_algorithmic_bytes": 4.0 * (M * K + K * N + M * N)}


if __name__ == "__main__":
    main()

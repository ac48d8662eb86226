"""Device engine: owns the HBM arenas (torch tensors as plain containers) and drives libskf.so.

Layout in HBM (all fp32 unless noted):
  params / grads / adam_m / adam_v : one flat arena each, layout from skf_param_table()
  workspace                        : every activation saved for backward + scratch, named tensors
The arithmetic of the hot path (models/sketchformer.py:325-349 of the reference) happens entirely in
the CUDA kernels behind include/skf.h; this file only moves pointers.
"""
import ctypes as C
import os
from typing import Dict

import torch

from . import _lib
from ._lib import SkfCfg, SkfParamEntry, check


def create_comm(world_size, rank, device):
    """NCCL communicator for the gradient exchange, through the C ABI (skf_comm_unique_id / skf_comm_create): rank 0
    draws the 128-byte id and the default torch.distributed group (any backend) carries it to the other ranks.
    Returns the opaque handle skf_plan_set_comm / skf_allreduce_bucket take."""
    import torch.distributed as dist
    lib = _lib.load()
    torch.cuda.set_device(device)
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        check(lib.skf_comm_unique_id(C.c_void_p(uid.data_ptr())), "skf_comm_unique_id")
    on_gpu = dist.get_backend() == "nccl"
    buf = uid.to(device) if on_gpu else uid
    dist.broadcast(buf, src=0)
    uid = buf.cpu().contiguous()
    comm = C.c_void_p()
    check(lib.skf_comm_create(C.c_void_p(uid.data_ptr()), int(world_size), int(rank), C.byref(comm)), "skf_comm_create")
    return comm.value


def make_cfg(B, T, d, f, H, L, V, n_classes, lowerdim, *, continuous=False, blind_decoder_mask=True,
             do_classification=True, do_reconstruction=True, class_weight=1.0, recon_weight=1.0,
             world_size=1, gemm_backend=0, attn_backend=0, gemm_precision=2, dropout_rate=0.0,
             dropout_seed=0, attn_version=1, class_buffer_layers=0, class_dropout=0.1) -> SkfCfg:
    return SkfCfg(B=B, T=T, d=d, f=f, H=H, L=L, V=V, C=n_classes, u=lowerdim, continuous=int(continuous),
                  blind_decoder_mask=int(blind_decoder_mask), do_classification=int(do_classification),
                  do_reconstruction=int(do_reconstruction), warmup_steps=5000, class_weight=class_weight,
                  recon_weight=recon_weight, adam_beta1=0.9, adam_beta2=0.98, adam_eps=1e-9,
                  world_size=world_size, gemm_backend=gemm_backend, attn_backend=attn_backend,
                  gemm_precision=gemm_precision, dropout_rate=float(dropout_rate), dropout_seed=int(dropout_seed),
                  attn_version=int(attn_version), class_buffer_layers=int(class_buffer_layers),
                  class_dropout=float(class_dropout))


def param_table(cfg: SkfCfg):
    """skf_param_table(): name -> (offset in floats, rows, cols) of the flat parameter / gradient arena, and its size.
    Pure host code (no device needed)."""
    lib = _lib.load()
    n = lib.skf_param_count(C.byref(cfg))
    if n <= 0:
        check(int(n) if n < 0 else -1, "skf_param_count")
    ent = (SkfParamEntry * 512)()
    cnt = lib.skf_param_table(C.byref(cfg), ent, 512)
    if cnt < 0:
        check(cnt, "skf_param_table")
    return {ent[i].name.decode(): (int(ent[i].offset), int(ent[i].rows), int(ent[i].cols)) for i in range(cnt)}, int(n)


def arena_name(name, d):
    """oracle / Keras variable name -> (arena tensor name, column slice or None): wq|wk|wv of a self-attention block
    are fused column-wise into wqkv, the cross-attention wk|wv into wkv (include/skf.h "parameter arena")."""
    for tag, fused, order in ((".mha.", "wqkv", ("wq", "wk", "wv")), (".mha1.", "wqkv", ("wq", "wk", "wv")),
                              (".mha2.", "wkv", ("wk", "wv"))):
        if tag in name:
            head, rest = name.split(tag)
            w, kind = rest.split(".")
            if w in order:
                i = order.index(w)
                return f"{head}{tag}{fused}.{kind}", slice(i * d, (i + 1) * d)
    return name, None


def fill_arena(arena, table, d, named):
    """Write tensors named as in oracle/sketchformer_oracle.py into a flat arena laid out by `table` (any device)."""
    for name, val in named.items():
        tname, sl = arena_name(name, d)
        off, r, c = table[tname]
        dst = arena[off: off + r * c].view(r, c)
        v = torch.as_tensor(val).detach().to(arena.dtype)
        if name == "pool.V" or v.dim() == 1:
            v = v.reshape(1, -1)
        if sl is None:
            dst.copy_(v)
        else:
            dst[:, sl].copy_(v)
    return arena


def grad_buckets(cfg: SkfCfg, n_buckets=4):
    """skf_grad_buckets(): [(lo, hi)] arena ranges in backward-completion order.  Pure host code."""
    lib = _lib.load()
    lo, hi = (C.c_int64 * 4)(), (C.c_int64 * 4)()
    n = lib.skf_grad_buckets(C.byref(cfg), int(n_buckets), lo, hi)
    if n < 0:
        check(n, "skf_grad_buckets")
    return [(int(lo[i]), int(hi[i])) for i in range(n)]


class Engine:
    METRIC_NAMES = ("recon_loss", "recon_acc", "class_loss", "class_acc", "total_loss")

    def __init__(self, cfg: SkfCfg, device="cuda:0", training=True, share_from=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.SkfError("no CUDA device: the Sketchformer engine has no CPU path")
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        if not self.lib.skf_device_ok():
            raise _lib.SkfError("device is not sm_100 (B200): libskf.so is built for sm_100a only")
        self.cfg = cfg
        n = self.lib.skf_param_count(C.byref(cfg))
        if n <= 0:
            check(int(n) if n < 0 else -1, "skf_param_count")
        self.n_params_arena = int(n)
        ent = (SkfParamEntry * 512)()
        cnt = self.lib.skf_param_table(C.byref(cfg), ent, 512)
        if cnt < 0:
            check(cnt, "skf_param_table")
        self.table = {ent[i].name.decode(): (int(ent[i].offset), int(ent[i].rows), int(ent[i].cols))
                      for i in range(cnt)}
        f32 = dict(dtype=torch.float32, device=self.device)
        if share_from is not None:
            # another plan (other batch size) over the SAME parameter / optimizer arenas
            assert share_from.table == self.table, "shared arenas need an identical parameter layout"
            self.params, self.grads = share_from.params, share_from.grads
            self.adam_m, self.adam_v = share_from.adam_m, share_from.adam_v
        else:
            self.params = torch.zeros(n, **f32)
            self.grads = torch.zeros(n, **f32) if training else None
            self.adam_m = torch.zeros(n, **f32) if training else None
            self.adam_v = torch.zeros(n, **f32) if training else None
        wsb = int(self.lib.skf_workspace_bytes(C.byref(cfg)))
        self.workspace = torch.empty(wsb, dtype=torch.uint8, device=self.device)
        self._plan = C.c_void_p()

        def ptr(t):
            return C.c_void_p(t.data_ptr()) if t is not None else None

        check(self.lib.skf_plan_create(C.byref(cfg), ptr(self.params), ptr(self.grads), ptr(self.adam_m),
                                       ptr(self.adam_v), ptr(self.workspace), wsb, C.byref(self._plan)),
              "skf_plan_create")
        check(self.lib.skf_plan_init(self._plan, self._stream()), "skf_plan_init")
        self.eager_train_steps = 0      # train steps this engine has run outside a CUDA graph (capture needs >= 1)

    def __del__(self):
        try:
            if getattr(self, "_plan", None) and self._plan.value:
                self.lib.skf_plan_destroy(self._plan)
                self._plan = C.c_void_p()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ---- workspace introspection -----------------------------------------------------------
    def tensor(self, name, shape=None, dtype=torch.float32):
        off, nb = C.c_size_t(), C.c_size_t()
        check(self.lib.skf_plan_tensor(self._plan, name.encode(), C.byref(off), C.byref(nb)), f"tensor {name}")
        t = self.workspace[off.value: off.value + nb.value].view(dtype)
        return t.view(*shape) if shape is not None else t

    # ---- parameters: Keras/oracle naming <-> arena ------------------------------------------------
    def _view(self, arena, name):
        off, r, c = self.table[name]
        return arena[off: off + r * c].view(r, c)

    def _map(self, name):
        """oracle name -> (arena tensor name, column slice or None)"""
        return arena_name(name, self.cfg.d)

    def load_params(self, named: Dict[str, torch.Tensor]):
        """Import tensors named as in oracle/sketchformer_oracle.py (Keras shapes, kernel=(in,out))."""
        self._import(self.params, named)

    def load_adam(self, m: Dict[str, torch.Tensor], v: Dict[str, torch.Tensor]):
        """Adam first / second moments, same naming and shapes as load_params()."""
        self._import(self.adam_m, m)
        self._import(self.adam_v, v)

    def _import(self, arena, named):
        seen = set()
        for name, val in named.items():
            tname, sl = self._map(name)
            dst = self._view(arena, tname)
            v = torch.as_tensor(val).detach().to(torch.float32)
            if name == "pool.V" or v.dim() == 1:
                v = v.reshape(1, -1)
            v = v.to(self.device)
            if sl is None:
                assert tuple(dst.shape) == tuple(v.shape), (name, tuple(dst.shape), tuple(v.shape))
                dst.copy_(v)
            else:
                dst[:, sl].copy_(v)
            seen.add(tname)
        missing = set(self.table) - seen
        if missing:
            raise KeyError(f"parameters not provided: {sorted(missing)[:5]}...")

    def _export(self, arena, like: Dict[str, torch.Tensor]):
        out = {}
        for name, ref in like.items():
            tname, sl = self._map(name)
            src = self._view(arena, tname)
            v = src if sl is None else src[:, sl]
            out[name] = v.detach().cpu().reshape(ref.shape).clone()
        return out

    def export_params(self, like):
        return self._export(self.params, like)

    def export_adam(self, like):
        return self._export(self.adam_m, like), self._export(self.adam_v, like)

    def export_grads(self, like):
        return self._export(self.grads, like)

    # ---- hot path ----------------------------------------------------------------------------
    def _check_tokens(self, tokens):
        """[B,T] int64 token ids, or [B,T,5] float32 stroke-5 rows when cfg.continuous (sketchformer.py:316-322)."""
        assert tokens.is_cuda and tokens.is_contiguous()
        if self.cfg.continuous:
            assert tokens.dtype == torch.float32 and tuple(tokens.shape) == (self.cfg.B, self.cfg.T, 5), tokens.shape
        else:
            assert tokens.dtype == torch.int64 and tuple(tokens.shape) == (self.cfg.B, self.cfg.T), tokens.shape

    def train_step(self, tokens: torch.Tensor, labels: torch.Tensor, apply_adam=True):
        self._check_tokens(tokens)
        assert labels.is_cuda and labels.dtype == torch.int64 and labels.numel() == self.cfg.B
        check(self.lib.skf_train_step(self._plan, C.c_void_p(tokens.data_ptr()), C.c_void_p(labels.data_ptr()),
                                      int(apply_adam), self._stream()), "skf_train_step")
        if not torch.cuda.is_current_stream_capturing():
            self.eager_train_steps += 1

    # ---- CUDA-graph replay of the whole step ------------------------------------------------------
    def capture_train_step(self, apply_adam=True):
        """Capture fwd + losses + bwd (+ Adam) into a CUDA graph that reads the static buffers
        `self.graph_tokens` / `self.graph_labels`.  The ~280 kernel launches of a step then cost one
        cudaGraphLaunch on the host.  Call after at least one eager step (function attributes, TMEM and
        tensor-map encoders are initialised lazily)."""
        if getattr(self, "_graph", None) is not None and self._graph_adam == bool(apply_adam):
            return self._graph
        B, T = self.cfg.B, self.cfg.T
        if self.cfg.continuous:
            self.graph_tokens = torch.zeros(B, T, 5, dtype=torch.float32, device=self.device)
        else:
            self.graph_tokens = torch.zeros(B, T, dtype=torch.int64, device=self.device)
        self.graph_labels = torch.zeros(B, 1, dtype=torch.int64, device=self.device)
        torch.cuda.synchronize(self.device)
        l0 = self.launches()
        g = torch.cuda.CUDAGraph()
        # captured on a high-priority stream: the plan's side stream (weight gradients, lowest priority) then only
        # takes SMs the critical path leaves idle
        prio = int(os.environ.get("SKF_GRAPH_PRIORITY", "-1"))
        with torch.cuda.graph(g, stream=torch.cuda.Stream(self.device, priority=prio)):
            self.train_step(self.graph_tokens, self.graph_labels, apply_adam)
        self.graph_launches = self.launches() - l0     # kernels per replay
        self._graph, self._graph_adam = g, bool(apply_adam)
        return g

    def replay_train_step(self, tokens=None, labels=None):
        """Copy a device batch into the static buffers (if given) and replay the captured step."""
        if tokens is not None:
            self.graph_tokens.copy_(tokens, non_blocking=True)
            self.graph_labels.copy_(labels.view(self.cfg.B, 1), non_blocking=True)
        self._graph.replay()

    def apply_adam(self):
        check(self.lib.skf_apply_adam(self._plan, self._stream()), "skf_apply_adam")

    def encode(self, tokens):
        self._check_tokens(tokens)
        check(self.lib.skf_encode(self._plan, C.c_void_p(tokens.data_ptr()), self._stream()), "skf_encode")

    def forward(self, tokens):
        self._check_tokens(tokens)
        check(self.lib.skf_forward(self._plan, C.c_void_p(tokens.data_ptr()), self._stream()), "skf_forward")

    def decode(self, embedding=None, expected_len=None, sos=1002, eos=1003, chunk=32):
        """Greedy decode (models/sketchformer.py:255-311) from `embedding` (device [B,d]; None = the one left
        in the workspace by encode()).  Steps are enqueued `chunk` at a time; between chunks ONE int32 is read
        back to stop early once every row has finished.  Returns the prefix incl. the start token:
        int32 [B, n] tokens or float32 [B, n, 5] strokes, n <= T+1 (the reference's `output`)."""
        B, T = self.cfg.B, self.cfg.T
        if embedding is not None:
            assert embedding.is_cuda and embedding.dtype == torch.float32 and embedding.is_contiguous()
            E = self.cfg.d if (self.cfg.attn_version in (0, 1) or not self.cfg.u) else self.cfg.u
            want = (B, E) if self.cfg.u else (B, T, E)       # lowerdim=0: the embedding is the encoder output
            assert tuple(embedding.shape) == want, tuple(embedding.shape)
        if expected_len is not None:
            expected_len = expected_len.to(device=self.device, dtype=torch.int32).contiguous()
            assert expected_len.numel() == B
        check(self.lib.skf_decode_begin(self._plan, C.c_void_p(embedding.data_ptr()) if embedding is not None else None,
                                        C.c_void_p(expected_len.data_ptr()) if expected_len is not None else None,
                                        int(sos), int(eos), self._stream()), "skf_decode_begin")
        if not self.cfg.do_reconstruction:
            return None                      # classification of the embedding only ("cls.argmax")
        state = self.tensor("gen.state", dtype=torch.int32)
        cols, i = 0, 0
        # The ~50 small launches of a decoder position are replayed as CUDA graphs, one per chunk of positions
        # (captured on the second decode of this plan with the same chunking: the first runs eagerly, which also
        # initialises function attributes and tensor-map encoders).  The reference re-runs the whole decoder on the
        # growing prefix and synchronises with the host at every token (models/sketchformer.py:274-301).
        graphs = getattr(self, "_decode_graphs", None)
        if graphs is None:
            graphs = self._decode_graphs = {"warm": False}
        use_graphs = graphs["warm"] and os.environ.get("SKF_DECODE_GRAPHS", "1") != "0"
        if not self.cfg.blind_decoder_mask and expected_len is None:
            # the uncached mode (include/skf.h): step i is i + 1 position steps — far too many launches for one graph
            # per chunk, and an early stop is worth checking for more often
            use_graphs, chunk = False, min(chunk, 8)
        while i < T and cols == 0:
            n = min(chunk, T - i)
            if use_graphs:
                g = graphs.get((i, n, int(sos), int(eos)))
                if g is None:
                    g = torch.cuda.CUDAGraph()
                    torch.cuda.synchronize(self.device)
                    with torch.cuda.graph(g, stream=self._graph_stream()):
                        check(self.lib.skf_decode_steps(self._plan, i, n, self._stream()), "skf_decode_steps")
                    graphs[(i, n, int(sos), int(eos))] = g
                g.replay()
            else:
                check(self.lib.skf_decode_steps(self._plan, i, n, self._stream()), "skf_decode_steps")
            i += n
            cols = int(state[0].item())
        graphs["warm"] = True
        cols = cols or T + 1
        if self.cfg.continuous:
            return self.tensor("gen.strokes", (B, T + 1, 5))[:, :cols].clone()
        return self.tensor("gen.tokens", (B, T + 1), dtype=torch.int64)[:, :cols].to(torch.int32)

    # ---- data parallel (include/skf.h "data-parallel exchange") -----------------------------------
    def set_comm(self, comm, n_buckets=4):
        """From now on train_step() all-reduces the gradient arena over the ranks of `comm` (an opaque handle from
        `create_comm`) in `n_buckets` buckets overlapped with the backward pass, before the optimizer.  A captured
        graph of the old step is dropped."""
        check(self.lib.skf_plan_set_comm(self._plan, C.c_void_p(comm) if comm else None, int(n_buckets)),
              "skf_plan_set_comm")
        self._graph = None

    def grad_buckets(self, n_buckets=4):
        return grad_buckets(self.cfg, n_buckets)

    def allreduce_grads(self, comm):
        """One in-place NCCL sum of the whole gradient arena on the current stream (tests, the dp check of bench.py)."""
        check(self.lib.skf_allreduce_bucket(C.c_void_p(self.grads.data_ptr()), self.grads.numel(), C.c_void_p(comm),
                                            self._stream()), "skf_allreduce_bucket")

    def _graph_stream(self):
        if getattr(self, "_gstream", None) is None:
            self._gstream = torch.cuda.Stream(self.device)
        return self._gstream

    def launches(self):
        return int(self.lib.skf_plan_launches(self._plan))

    def set_step(self, n: int):
        """optimizer.iterations (drives the Noam schedule and Adam's bias correction); restored on resume."""
        self.tensor("step", dtype=torch.int64).fill_(int(n))

    def step_count(self):
        return int(self.tensor("step", dtype=torch.int64)[0].item())

    def merge_metrics_into(self, main):
        """Fold this plan's running sums into `main`'s and clear them (device-side, no synchronisation): the
        reference has ONE MetricManager whatever the batch size of a step (builders/keras_metrics.py:13-42)."""
        if main is self:
            return
        mine, theirs = self.tensor("stats_run", dtype=torch.float64), main.tensor("stats_run", dtype=torch.float64)
        theirs[:8] += mine[:8]
        theirs[8:13] = mine[8:13]
        mine.zero_()

    def reset_metrics(self):
        """MetricManager.reset (builders/keras_metrics.py): zero the running sums."""
        self.tensor("stats_run", dtype=torch.float64).zero_()

    def stats(self):
        """Raw running sums (device -> host read of 16 doubles)."""
        return self.tensor("stats_run", dtype=torch.float64).cpu().numpy().copy()

    def metrics(self):
        """Running means exactly like builders/keras_metrics.py:41-42 (never reset in train())."""
        s = self.stats()
        steps = max(s[7], 1.0)
        return {"recon_loss": s[0] / steps, "recon_acc": s[1] / max(s[2], 1.0), "class_loss": s[3] / steps,
                "class_acc": s[4] / max(s[5], 1.0), "total_loss": s[6] / steps}

    def last_step_metrics(self):
        s = self.stats()
        return dict(zip(self.METRIC_NAMES, (float(x) for x in s[8:13])))

"""Host-logic test double for sketchformer_b200.engine.Engine (test infrastructure, never imported by the product).

The reference's control flow around the hot path — BaseModel.train(), the checkpoint cadence, compute_metrics_from —
can only be executed side by side with the plugin in THIS container (the reference tree is here, a B200 is not), so the
plugin's host code (sketchformer_b200/model.py) is run on an engine whose arithmetic is a few cheap deterministic
functions of its inputs.  The attribute and call protocol is the real Engine's; the parameter layout is the real one
(skf_param_table through the C ABI, which needs no device)."""
import numpy as np
import torch

from sketchformer_b200.engine import Engine, fill_arena, param_table, arena_name


class FakeEngine:
    METRIC_NAMES = Engine.METRIC_NAMES
    instances = []

    def __init__(self, cfg, device="cpu", training=True, share_from=None):
        self.cfg, self.device = cfg, torch.device("cpu")
        self.table, n = param_table(cfg)
        if share_from is not None:
            assert share_from.table == self.table
            self.params, self.grads, self.adam_m, self.adam_v = (share_from.params, share_from.grads, share_from.adam_m,
                                                                 share_from.adam_v)
            self._step = share_from._step
        else:
            self.params = torch.zeros(n)
            self.grads = torch.zeros(n) if training else None
            self.adam_m = torch.zeros(n) if training else None
            self.adam_v = torch.zeros(n) if training else None
            self._step = torch.zeros(1, dtype=torch.int64)
        self._stats = torch.zeros(16, dtype=torch.float64)
        self._ws = {}
        self.eager_train_steps = 0
        self.calls = []
        self._graph = None
        FakeEngine.instances.append(self)

    # ---- parameters ------------------------------------------------------------------------------------------
    def load_params(self, named):
        fill_arena(self.params, self.table, self.cfg.d, named)

    def load_adam(self, m, v):
        fill_arena(self.adam_m, self.table, self.cfg.d, m)
        fill_arena(self.adam_v, self.table, self.cfg.d, v)

    def _export(self, arena, like):
        out = {}
        for name, ref in like.items():
            tname, sl = arena_name(name, self.cfg.d)
            off, r, c = self.table[tname]
            src = arena[off: off + r * c].view(r, c)
            out[name] = (src if sl is None else src[:, sl]).reshape(ref.shape).clone()
        return out

    def export_params(self, like):
        return self._export(self.params, like)

    def export_adam(self, like):
        return self._export(self.adam_m, like), self._export(self.adam_v, like)

    # ---- train step: losses are functions of the batch, the "optimizer" nudges the arena ---------------------------
    def train_step(self, tokens, labels, apply_adam=True, _eager=True):
        B = int(tokens.shape[0])
        assert B == self.cfg.B and labels.numel() == B
        t = tokens.double().reshape(B, -1)
        recon_loss = float(t.mean() % 7.0) + 1.0 / (1 + int(self._step))
        class_loss = float(labels.double().mean() % 3.0) + 0.5
        n_tok = float((t != 0).sum())
        hit_tok = float(((t.long() % 2) == 0).sum())
        hit_cls = float((labels.reshape(-1) % 2 == 0).sum())
        total = self.cfg.recon_weight * recon_loss + self.cfg.class_weight * class_loss
        s = self._stats
        s[0] += recon_loss; s[1] += hit_tok; s[2] += n_tok; s[3] += class_loss; s[4] += hit_cls; s[5] += B
        s[6] += total; s[7] += 1
        s[8:13] = torch.tensor([recon_loss, hit_tok / max(n_tok, 1), class_loss, hit_cls / B, total], dtype=torch.float64)
        # a "gradient" that is a batch mean of a per-sample quantity, like the real losses' gradients: per-rank
        # gradients summed over the ranks times 1/world equal the full-batch gradient (tests/test_dp_gloo.py)
        self.grads.fill_(float(t.sum(1).mean()) * 1e-3)
        if apply_adam:
            self.apply_adam()
        if _eager:
            self.eager_train_steps += 1
        self.calls.append(("train_step", B, bool(apply_adam), bool(_eager)))

    def apply_adam(self):
        self.params.add_(self.grads, alpha=-1e-3 / max(1, self.cfg.world_size))     # 1/world folded into the optimizer
        self.adam_m.mul_(0.9).add_(self.grads, alpha=0.1)
        self.adam_v.mul_(0.98).add_(self.grads * self.grads, alpha=0.02)
        self._step += 1

    def capture_train_step(self, apply_adam=True):
        if self._graph is None or self._graph != bool(apply_adam):
            B, T = self.cfg.B, self.cfg.T
            self.graph_tokens = torch.zeros((B, T, 5) if self.cfg.continuous else (B, T),
                                            dtype=torch.float32 if self.cfg.continuous else torch.int64)
            self.graph_labels = torch.zeros(B, 1, dtype=torch.int64)
            self._graph = bool(apply_adam)
            self.calls.append(("capture", bool(apply_adam)))
        return self._graph

    def replay_train_step(self, tokens=None, labels=None):
        self.train_step(self.graph_tokens, self.graph_labels, self._graph, _eager=False)

    def set_comm(self, comm, n_buckets=4):
        self._graph = None

    # ---- metrics (same slots as the device's stats_run) -------------------------------------------------------------
    def tensor(self, name, shape=None, dtype=torch.float32):
        if name == "stats_run":
            return self._stats
        if name == "step":
            return self._step
        t = self._ws[name]
        return t.view(*shape) if shape is not None else t

    def stats(self):
        return self._stats.numpy().copy()

    metrics = Engine.metrics
    last_step_metrics = Engine.last_step_metrics
    merge_metrics_into = Engine.merge_metrics_into
    reset_metrics = Engine.reset_metrics

    def set_step(self, n):
        self._step.fill_(int(n))

    def step_count(self):
        return int(self._step[0])

    # ---- inference: class = second token mod C, embedding = the first E tokens, reconstruction = echo -------------
    def encode(self, tokens):
        B = int(tokens.shape[0])
        x = tokens.reshape(B, -1).float()
        E = self.cfg.d if (self.cfg.attn_version in (0, 1) or not self.cfg.u) else self.cfg.u
        emb = torch.zeros(B, E)
        w = min(E, x.shape[1])
        emb[:, :w] = x[:, :w]
        self._ws["embedding"] = emb
        am = (x[:, 1].long() % self.cfg.C).to(torch.int32)
        self._ws["cls.argmax"] = am
        self._ws["cls.probs"] = torch.nn.functional.one_hot(am.long(), self.cfg.C).float()
        self._last_tokens = tokens
        self.calls.append(("encode", B))

    def decode(self, embedding=None, expected_len=None, sos=1002, eos=1003, chunk=32):
        B = self.cfg.B
        self.calls.append(("decode", B))
        if embedding is not None:
            am = (embedding[:, 1].long() % self.cfg.C).to(torch.int32)
            self._ws["cls.argmax"] = am
            body = embedding[:, : self.cfg.T].long() % 1000 + 1
        else:
            body = self._last_tokens.reshape(B, -1)[:, : self.cfg.T].long()
        return torch.cat([torch.full((B, 1), int(sos), dtype=torch.int64), body], 1).to(torch.int32)


def install(monkeypatch=None):
    """Run sketchformer_b200.model.Transformer on FakeEngine: the engine class and the two pinned-memory staging helpers
    (which need a CUDA context) are the only things replaced."""
    from sketchformer_b200 import model as M
    FakeEngine.instances = []
    if monkeypatch is None:          # a spawned worker process: nothing to restore afterwards
        class _Set:
            @staticmethod
            def setattr(obj, name, value):
                setattr(obj, name, value)
        monkeypatch = _Set
    monkeypatch.setattr(M, "Engine", FakeEngine)

    def to_device(self, key, arr, dtype):
        t = torch.as_tensor(np.ascontiguousarray(arr)) if not torch.is_tensor(arr) else arr
        return t.to(dtype).contiguous()

    def h2d_into(self, key, arr, dtype, dst):
        t = torch.as_tensor(np.ascontiguousarray(arr)) if not torch.is_tensor(arr) else arr
        dst.copy_(t.to(dtype).view(dst.shape))

    monkeypatch.setattr(M.Transformer, "_to_device", to_device)
    monkeypatch.setattr(M.Transformer, "_h2d_into", h2d_into)
    return FakeEngine

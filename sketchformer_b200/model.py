"""Host-side mirror of the reference's model plugin for the hot path.

`Transformer` presents the surface of models/sketchformer.py:17-365 + core/models.py:18-136 that
train.py / evaluate-metrics.py / the metrics and experiments drive: same registry
name, hparam names/defaults, constructor, `train_on_batch`, `predict_class`, `encode_from_seq`,
`predict`, `predict_from_embedding`.  Every byte of arithmetic runs in libskf.so (include/skf.h);
this class only stages host batches into HBM and reads results back.  There is no CPU path: without
a B200 and the built library the constructor raises.
"""
import os
import re

import numpy as np
import torch

from . import evaluation, tf_checkpoint
from .engine import Engine, make_cfg
from .hparams import HParams, combine_hparams_into_one
from .init import init_params


_ROLLING = re.compile(r"^ckpt-(\d+)\.pt$")


class Transformer:
    name = "sketch-transformer-tf2"                                   # models/sketchformer.py:18
    quick_metrics = ["recon_loss", "recon_acc", "class_loss", "class_acc", "total_loss"]
    slow_metrics = ["sketch-reconstruction", "val-clas-acc", "tsne", "tsne-predicted"]
    MAX_EXTRA_PLANS = 4      # plans (workspaces) kept for batch sizes other than hps['batch_size']

    # ---- hparams (names, types and defaults are the reference's) ---------------------------
    @classmethod
    def base_default_hparams(cls):
        """core/models.py:18-31"""
        return HParams(batch_size=128, num_epochs=10, save_every=1., safety_save=.5, autograph=True,
                       log_every=100, notify_every=1000, slack_config="token.secret", goal="No description")

    @classmethod
    def specific_default_hparams(cls):
        """models/sketchformer.py:24-53"""
        return HParams(num_layers=4, d_model=128, dff=512, num_heads=8, dropout_rate=0.1,
                       lowerdim=256, attn_version=1,
                       do_classification=True, class_weight=1.0, class_buffer_layers=0, class_dropout=0.1,
                       do_reconstruction=True, recon_weight=1.0, blind_decoder_mask=True,
                       is_training=True, optimizer="Adam", lr=0.01, lr_scheduler="WarmupDecay",
                       warmup_steps=10000)

    @classmethod
    def default_hparams(cls):
        """core/models.py:109-114"""
        return combine_hparams_into_one(cls.specific_default_hparams(), cls.base_default_hparams())

    @classmethod
    def get_config_filepath(cls, output_dir, exp_id):
        """core/models.py:116-119"""
        return os.path.join(output_dir, "{}-{}".format(cls.name, exp_id), "config.json")

    @classmethod
    def parse_hparams(cls, base, specific):
        """core/models.py:121-136"""
        hps = cls.default_hparams()
        if base is not None:
            hps = hps.parse(base)
        if specific is not None:
            hps = hps.parse(specific)
        return hps

    # ---- construction ----------------------------------------------------------------------
    def __init__(self, hps, dataset, out_dir=None, experiment_id="0", device="cuda:0", world_size=1, seed=0,
                 gemm_precision=None, use_cuda_graph=True, rank=0, dp_buckets=None):
        self.hps = hps if isinstance(hps, dict) else dict(hps.values())
        self.dataset = dataset
        self.experiment_id = experiment_id
        dhps = dataset.hps if isinstance(dataset.hps, dict) else dict(dataset.hps.values())
        self.continuous = bool(dhps["use_continuous_data"])
        self.vocab_size = None if self.continuous else dataset.tokenizer.VOCAB_SIZE   # sketchformer.py:59
        self.seq_len = dhps["max_seq_len"]                                            # sketchformer.py:60
        # core/models.py:60: n_samples // batch_size.  Under data parallelism a step consumes world_size batches
        # (one per rank), so an epoch is world_size times fewer steps and the checkpoint cadence follows it.
        self.batches_per_epoch = max(1, getattr(dataset, "n_samples", 0) // (self.hps["batch_size"] * max(1, world_size)))
        if out_dir is not None:                                                       # core/models.py:68-81
            self.out_dir = os.path.join(out_dir, self.identifier)
            self.plots_out_dir = os.path.join(self.out_dir, "plots")
            self.wgt_out_dir = os.path.join(self.out_dir, "weights")
            self.tmp_out_dir = os.path.join(self.out_dir, "tmp")
            for p in (self.out_dir, self.plots_out_dir, self.wgt_out_dir, self.tmp_out_dir):
                os.makedirs(p, exist_ok=True)
            self.config_filepath = self.get_config_filepath(out_dir, experiment_id)
        self.current_step = 0
        self._device = torch.device(device)
        self._world = world_size
        self._rank = rank
        self._seed = seed
        self._precision = gemm_precision
        self._use_graph = use_cuda_graph
        self._engines = {}
        self._stage = {}
        # data parallel: dp_buckets >= 1 -> bucketed NCCL all-reduce inside the step, overlapped with the backward pass
        # and captured in the CUDA graph (skf_plan_set_comm); 0 -> one torch.distributed all_reduce after the step
        # Default: the in-step exchange on 2 ranks, the torch path on more.  Both run on 2 and 8 ranks and measure the same
        # step time (profiles/r2_dp2_*, r2_dp8_*, r2_f2_dp8_weak_instep.json); the in-step path
        # has not been run on 4 ranks, so it stays opt-in there.
        if dp_buckets is None:
            dp_buckets = 4 if world_size == 2 else 0
        self._dp_buckets = int(dp_buckets) if world_size > 1 else 0
        self._comm = None
        self._dp_eager_steps = 0
        self.build_model()

    @property
    def embedding_dim(self):
        """SelfAttnV1 pools to d_model; SelfAttnV2 adds Dense(lowerdim) (transformer.py:129-130)."""
        return self.hps["d_model"] if (self.hps["attn_version"] == 1 or not self.hps["lowerdim"]) else self.hps["lowerdim"]

    @property
    def identifier(self):
        return "{}-{}".format(self.name, self.experiment_id)

    def _unsupported(self):
        h = self.hps
        if not h["lowerdim"] and (h["do_classification"] or not h["do_reconstruction"]):
            return ("lowerdim=0 with do_classification=True (the reference builds no classify_layer then, "
                    "models/sketchformer.py:96-107)")
        if h["optimizer"].lower() != "adam":
            return "optimizer=%s (only Adam is built)" % h["optimizer"]
        if h["lr_scheduler"] not in ("WarmupDecay", "warmup-decay"):
            return "lr_scheduler=%s (the reference's step-decay call site raises too)" % h["lr_scheduler"]
        return None

    def build_model(self):
        """models/sketchformer.py:63-129 — here: validate the flags and create the training plan."""
        why = self._unsupported()
        if why:
            raise NotImplementedError(f"sketch-transformer-tf2 on B200: {why} is not built yet")
        self._weights = None
        self._engine(self.hps["batch_size"], training=True)

    def _cfg(self, B):
        h = self.hps
        kw = {} if self._precision is None else {"gemm_precision": self._precision}
        return make_cfg(B, self.seq_len, h["d_model"], h["dff"], h["num_heads"], h["num_layers"],
                        self.vocab_size or 0, self.dataset.n_classes, h["lowerdim"], continuous=self.continuous,
                        blind_decoder_mask=h["blind_decoder_mask"], do_classification=h["do_classification"],
                        do_reconstruction=h["do_reconstruction"], class_weight=h["class_weight"],
                        recon_weight=h["recon_weight"], world_size=self._world, dropout_rate=h["dropout_rate"], attn_version=h["attn_version"],
                        class_buffer_layers=h["class_buffer_layers"], class_dropout=h["class_dropout"],
                        dropout_seed=(0x5EED + 7919 * self._seed + 104729 * self._rank) & 0xFFFFFFFF, **kw)

    def _layout_hps(self):
        """model + data hparams that decide which variables exist and their shapes (init.py, tf_checkpoint.key_map)"""
        return dict(self.hps, max_seq_len=self.seq_len, vocab_size=self.vocab_size, n_classes=self.dataset.n_classes,
                    use_continuous_data=self.continuous)

    def _engine(self, B, training=False):
        """One plan per batch size; all plans share the parameter arena of the training plan."""
        if B in self._engines:
            eng = self._engines.pop(B)
            self._engines[B] = eng            # most recently used last
            return eng
        first = not self._engines
        # batch_iterator cuts the running batch short at every end of split and the cut size varies with the chunk
        # order: keep the training plan plus the few most recently used other sizes, not one workspace per size
        others = [b for b, e in self._engines.items() if e is not self._train_engine]
        while len(others) >= self.MAX_EXTRA_PLANS:
            del self._engines[others.pop(0)]
        eng = Engine(self._cfg(B), device=self._device, training=training or first,
                     share_from=None if first else self._train_engine)
        if first:
            self._train_engine = eng
            eng.load_params(init_params(self._layout_hps(), seed=self._seed))
        if self._dp_buckets and eng.grads is not None:
            if self._comm is None:
                from .engine import create_comm
                self._comm = create_comm(self._world, self._rank, self._device)
            eng.set_comm(self._comm, self._dp_buckets)
        self._engines[B] = eng
        return eng

    # ---- staging ---------------------------------------------------------------------------
    def _to_device(self, key, arr, dtype):
        """Host numpy (or torch CPU) -> pinned staging ring -> HBM, asynchronously on the current stream.
        A slot is reused only after the copy that last read it has completed (event)."""
        t = torch.as_tensor(np.ascontiguousarray(arr)) if not torch.is_tensor(arr) else arr
        if t.is_cuda:
            return t.to(dtype).contiguous()
        t = t.to(dtype)
        ring = self._stage.setdefault((key, tuple(t.shape)), {"slots": [], "next": 0})
        if len(ring["slots"]) < 3:
            ring["slots"].append((torch.empty(t.shape, dtype=dtype).pin_memory(),
                                  torch.empty(t.shape, dtype=dtype, device=self._device), torch.cuda.Event()))
            host, dev, ev = ring["slots"][-1]
        else:
            host, dev, ev = ring["slots"][ring["next"]]
            ring["next"] = (ring["next"] + 1) % len(ring["slots"])
            ev.synchronize()
        host.copy_(t)
        dev.copy_(host, non_blocking=True)
        ev.record()
        return dev

    def _h2d_into(self, key, arr, dtype, dst):
        """Host batch -> pinned ring slot -> the graph's static device buffer `dst`."""
        t = torch.as_tensor(np.ascontiguousarray(arr)) if not torch.is_tensor(arr) else arr
        if t.is_cuda:
            dst.copy_(t.to(dtype).view(dst.shape), non_blocking=True)
            return
        t = t.to(dtype).view(dst.shape)
        ring = self._stage.setdefault(("g" + key, tuple(t.shape)), {"slots": [], "next": 0})
        if len(ring["slots"]) < 3:
            ring["slots"].append((torch.empty(t.shape, dtype=dtype).pin_memory(), torch.cuda.Event()))
            host, ev = ring["slots"][-1]
        else:
            host, ev = ring["slots"][ring["next"]]
            ring["next"] = (ring["next"] + 1) % len(ring["slots"])
            ev.synchronize()
        host.copy_(t)
        dst.copy_(host, non_blocking=True)
        ev.record()

    # ---- training --------------------------------------------------------------------------
    def train_on_batch(self, batch, sync_metrics=True):
        """models/sketchformer.py:351-359: `inp = tar = data`; returns the running-mean quick metrics
        (builders/keras_metrics.py:41-42).  `sync_metrics=False` skips the device->host read."""
        data, labels = batch
        B = int(data.shape[0])
        eng = self._engine(B, training=True)
        dt = torch.int64 if not self.continuous else torch.float32
        # the optimizer runs inside the step unless the gradients still have to be all-reduced by torch afterwards
        solo = self._world == 1 or self._dp_buckets > 0
        # Under data parallelism with the exchange inside the step, the step at which a rank starts capturing must be the
        # SAME on every rank: NCCL does host-side work of its own at capture time (buffer registration for the graph),
        # which is collective.  So the rule counts the eager steps THIS method has run on the training plan — every
        # rank calls it equally often — not the engine's own history (a rank whose plan also served some other eager
        # step, e.g. rank 0 in bench.py's gradient check, would otherwise capture one call early while the others run
        # the collective eagerly).
        eager_done = eng.eager_train_steps if not (self._world > 1 and self._dp_buckets) else self._dp_eager_steps
        if self._use_graph and eng is self._train_engine and eager_done >= 1:
            # (one eager step first — also right after restore_checkpoint_if_exists(), when current_step is already
            # large: function attributes and tensor-map encoders are initialised lazily, engine.py capture_train_step)
            # steady state: pinned host -> static device buffers -> one cudaGraphLaunch for the whole step
            eng.capture_train_step(apply_adam=solo)
            self._h2d_into("x", data, dt, eng.graph_tokens)
            self._h2d_into("y", labels, torch.int64, eng.graph_labels)
            eng.replay_train_step()
        else:
            x = self._to_device("x", data, dt)
            y = self._to_device("y", labels, torch.int64)
            if eng is not self._train_engine:          # odd batch size: same arenas, carry the optimizer step over
                eng.set_step(self._train_engine.step_count())
            eng.train_step(x, y, apply_adam=solo)
            if eng is self._train_engine:
                self._dp_eager_steps += 1
            if eng is not self._train_engine and solo:
                self._train_engine.set_step(eng.step_count())
        # one running mean over all steps, whatever plan ran them
        eng.merge_metrics_into(self._train_engine)
        if not solo:
            self._allreduce_and_step(eng)
        self.current_step += 1
        return self._quick_metrics(self._train_engine) if sync_metrics else None

    def _quick_metrics(self, eng):
        """Only the metrics build_model() registers (models/sketchformer.py:85-107): no recon_acc with
        continuous data, no class_* without the classifier, no recon_* without the decoder."""
        m = eng.metrics()
        if self.continuous or not self.hps["do_reconstruction"]:
            m.pop("recon_acc", None)
        if not self.hps["do_reconstruction"]:
            m.pop("recon_loss", None)
        if not self.hps["do_classification"]:
            m.pop("class_loss", None), m.pop("class_acc", None)
        return m

    def _allreduce_and_step(self, eng):
        """sum the flat gradient arena over the ranks (NCCL over NVLink; 1/world is folded
        into the Adam kernel), then the optimizer.  The reference has no multi-GPU path."""
        import torch.distributed as dist
        dist.all_reduce(eng.grads)
        eng.apply_adam()
        if eng is not self._train_engine:
            self._train_engine.set_step(eng.step_count())

    def prepare_for_start_of_epoch(self):
        pass

    def prepare_for_end_of_epoch(self):
        """models/sketchformer.py:364-365: metrics_manager.reset()"""
        for e in self._engines.values():
            e.reset_metrics()

    # ---- inference -------------------------------------------------------------------------
    def encode_from_seq(self, inp_seq):
        """models/sketchformer.py:162-168 -> {'embedding' (B,d) float32, 'class' (B,C) probabilities}"""
        inp_seq = np.asarray(inp_seq)
        if inp_seq.ndim == (2 if self.continuous else 1):
            inp_seq = inp_seq[None]
        B = inp_seq.shape[0]
        eng = self._engine(B)
        eng.encode(self._to_device("x", inp_seq, torch.float32 if self.continuous else torch.int64))
        if not self.hps["lowerdim"]:      # embedding = enc_output (B, T, d), models/sketchformer.py:157-158
            name = "enc.l%d.out2" % (self.hps["num_layers"] - 1)
            return {"embedding": eng.tensor(name, (B, self.seq_len, self.hps["d_model"])).cpu().numpy(), "class": None}
        out = {"embedding": eng.tensor("embedding", (B, self.embedding_dim)).cpu().numpy(), "class": None}
        if self.hps["do_classification"]:
            out["class"] = eng.tensor("cls.probs", (B, self.dataset.n_classes)).cpu().numpy()
            out["_argmax"] = eng.tensor("cls.argmax", dtype=torch.int32)[:B].cpu().numpy()
        return out

    def predict_class(self, inp_seq):
        """models/sketchformer.py:223-228 -> {'class' (B,) int32, 'embedding' (B,d)}"""
        out = self.encode_from_seq(inp_seq)
        am = out.pop("_argmax", None)
        if self.hps["do_classification"]:
            out["class"] = am.astype(np.int32)
        return out

    def _sos_eos(self):
        tok = getattr(self.dataset, "tokenizer", None)
        return (tok.SOS, tok.EOS) if tok is not None else (0, 0)

    def predict_from_embedding(self, emb, expected_len=None):
        """models/sketchformer.py:255-311 -> {'recon' (B, n<=T+1) int32 tokens | (B, n, 5) float32 strokes,
        'class' (B,) int32, 'attn_weights': None}.  KV-cached greedy decode on the device (skf_decode_*);
        `expected_len` is ignored when blind_decoder_mask=True, exactly as the reference's decode() does; with
        blind_decoder_mask=False and no expected_len the reference's growing `i + 1` cross mask is reproduced by
        recomputing the prefix at every step (slow, like the reference: experiments/interpolations_for_mturk.py:78).
        The per-step attention maps the reference returns are never materialised here."""
        emb = np.asarray(emb, dtype=np.float32)
        if emb.ndim == (1 if self.hps["lowerdim"] else 2):
            emb = emb[None]
        B = emb.shape[0]
        eng = self._engine(B)
        out = {}
        e = self._to_device("emb", emb, torch.float32)
        el = None
        if expected_len is not None and not self.hps["blind_decoder_mask"]:
            el = torch.as_tensor(np.asarray(expected_len)).to(torch.int32)
        sos, eos = self._sos_eos()
        recon = eng.decode(e, el, sos, eos)          # also classifies the embedding on the device
        if self.hps["do_reconstruction"]:
            out["recon"] = recon.cpu().numpy()
            out["attn_weights"] = None
        if self.hps["do_classification"]:
            out["class"] = eng.tensor("cls.argmax", dtype=torch.int32)[:B].cpu().numpy().astype(np.int32)
        return out

    def predict(self, inp_seq):
        """models/sketchformer.py:201-221 -> {'embedding', 'class' (B,) int32, 'recon', 'attn_weights'}"""
        inp_seq = np.asarray(inp_seq)
        out = self.predict_class(inp_seq)
        if self.hps["do_reconstruction"]:
            if inp_seq.ndim == (2 if self.continuous else 1):
                inp_seq = inp_seq[None]
            B = inp_seq.shape[0]
            eng = self._engine(B)
            tlen = None
            if not self.hps["blind_decoder_mask"]:          # exact length of the input sketch
                tlen = (inp_seq[..., -1] != 1).sum(-1) if self.continuous else (inp_seq > 0).sum(-1)
                tlen = torch.as_tensor(tlen).to(torch.int32)
            sos, eos = self._sos_eos()
            out["recon"] = eng.decode(None, tlen, sos, eos).cpu().numpy()   # embedding still in the workspace
            out["attn_weights"] = None
        return out

    # ---- training loop, reports, checkpoint cadence (core/models.py:163-223, 321-358) -----------------
    def train(self, max_steps=None):
        """core/models.py:163-197: total = batches_per_epoch * num_epochs steps from current_step; per step one
        batch from dataset.batch_iterator('train', batch_size), train_on_batch, status report, checkpoint
        cadence.  The slow-metric / Slack machinery of the reference is out of scope."""
        total_steps = self.batches_per_epoch * self.hps["num_epochs"]
        self.epoch = self.current_step // self.batches_per_epoch
        it = self.dataset.batch_iterator("train", self.hps["batch_size"], stop_at_end_of_split=False)
        todo = total_steps - self.current_step
        if max_steps is not None:
            todo = min(todo, max_steps)
        self.quick_metrics_history = getattr(self, "quick_metrics_history", {k: [] for k in self.quick_metrics})
        for _ in range(max(0, todo)):
            # the reference increments current_step first (core/models.py:184); train_on_batch does it here
            log_now = ((self.current_step + 1) % self.batches_per_epoch) % self.hps["log_every"] == 0
            qm = self.train_on_batch(next(it), sync_metrics=True)
            for k, v in qm.items():
                self.quick_metrics_history.setdefault(k, []).append(v)
            self.last_quick_metrics = qm
            if log_now:
                print(self.status_report())
            self.save_checkpoint_if_its_time()
            if self.current_step // self.batches_per_epoch > self.epoch:
                self.epoch = self.current_step // self.batches_per_epoch
                print(self.status_report())

    def status_report(self, end_of_epoch=False):
        """core/models.py:203-223 (log string only)."""
        cur_iter = self.current_step % self.batches_per_epoch
        s = "Epoch {} Batch {}/{}".format(getattr(self, "epoch", 0), cur_iter, self.batches_per_epoch)
        for k, v in getattr(self, "last_quick_metrics", {}).items():
            s = "{}|{}={:4.4f}".format(s, k, v)
        return s

    def save_checkpoint_if_its_time(self):
        """core/models.py:346-358: rolling 'safety' checkpoints (the 2 newest are kept, like
        CheckpointManager(max_to_keep=2)) every safety_save epochs and permanent ones every save_every."""
        if getattr(self, "wgt_out_dir", None) is None:
            return
        cur = self.current_step
        safety = max(1, int(self.hps["safety_save"] * self.batches_per_epoch))
        every = max(1, int(self.hps["save_every"] * self.batches_per_epoch))
        if (cur + 1) % safety == 0:
            self.save_checkpoint(os.path.join(self.wgt_out_dir, "ckpt-{}.pt".format(cur)))
            # rotate only the files this method wrote: the reference's CheckpointManager keeps `ckpt-N.index` /
            # `ckpt-N.data-*` tensor bundles in the same directory (core/models.py:326-331) and they are not ours
            mine = [(int(m.group(1)), f) for f in os.listdir(self.wgt_out_dir)
                    for m in [_ROLLING.match(f)] if m]
            for _, f in sorted(mine)[:-2]:
                os.remove(os.path.join(self.wgt_out_dir, f))
        if (cur + 1) % every == 0:
            self.save_checkpoint(os.path.join(self.wgt_out_dir, "step{}.pt".format(cur)))

    def save_checkpoint(self, path):
        torch.save(self.state_dict(), path)
        return path

    def restore_checkpoint_if_exists(self, checkpoint):
        """core/models.py:333-344: None -> nothing, 'latest' -> newest rolling checkpoint in weights/, else a path."""
        if checkpoint is None:
            return False
        if checkpoint == "latest":
            d = getattr(self, "wgt_out_dir", None)
            cands = sorted((f for f in (os.listdir(d) if d else []) if f.endswith(".pt")),
                           key=lambda f: os.path.getmtime(os.path.join(d, f)))
            tf_prefix = tf_checkpoint.latest_checkpoint(d) if d and not cands else None
            if not cands and tf_prefix is None:
                print("[Checkpoint] Not found")
                return False
            checkpoint = os.path.join(d, cands[-1]) if cands else tf_prefix
        if os.path.exists(checkpoint + ".index"):       # a TensorFlow tensor bundle written by the reference
            self.load_tf_checkpoint(checkpoint)
            print("[Checkpoint] Restored, step #{}".format(self.current_step))
            return True
        # tensors, ints and dicts / tuples of primitives only: the safe loader (no arbitrary unpickling)
        self.load_state_dict(torch.load(checkpoint, map_location="cpu", weights_only=True))
        print("[Checkpoint] Restored, step #{}".format(self.current_step))
        return True

    # ---- prediction sweeps for the slow metrics (models/evaluation_mixin.py:10-53) ----------------------
    def compute_predictions_on_validation_set(self):
        return evaluation.compute_predictions_on_set(self, "valid")

    def compute_predictions_on_test_set(self):
        return evaluation.compute_predictions_on_set(self, "test")

    def compute_predictions_on_set(self, set_type):
        return evaluation.compute_predictions_on_set(self, set_type)

    # ---- slow-metric plumbing of evaluate-metrics.py:53-61 (core/models.py:249-320) ----------------------
    def compute_metrics_from(self, chosen_metrics):
        """core/models.py:254-257: for every metric object gather the data its `input_type` names and run its
        `computation_worker` on it (synchronously).  The metric objects are the caller's: the reference's own
        `metrics.build_metric_by_name(...)` instances work unchanged (tests/test_integration_stub.py)."""
        for _, metric in chosen_metrics.items():
            input_data = self.gather_data_for_metric(metric.input_type)
            metric.computation_worker(input_data)

    def gather_data_for_metric(self, data_type):
        """core/models.py:280-292: `compute_<data_type>()` once per time window, cached."""
        data = self.load_precomputed_features_if_they_exist(data_type)
        if not data:
            try:
                gather_func = getattr(self, "compute_{}".format(data_type))
            except AttributeError:
                raise AttributeError("One of your metrics requires a compute_{} method".format(data_type))
            data = gather_func()
            self.save_precomputed_features(data_type, data)
            data = self.load_precomputed_features_if_they_exist(data_type)
        return data

    @staticmethod
    def _time_id_str():
        """utils/helpers.py:23-30: DDHHM with the minutes cut to the tens — the reference's cache window."""
        import datetime
        now = datetime.datetime.now()
        return "{:02d}{:02d}{:02d}".format(now.day, now.hour, now.minute)[:-1]

    def load_precomputed_features_if_they_exist(self, feature_type):
        """core/models.py:294-308.  The reference re-reads `tmp/precomputed_<type>_<time id>.pickle`; here the tuple
        stays in memory for the same ten-minute window and the file is only written (nothing is unpickled)."""
        hit = getattr(self, "_precomputed", {}).get(feature_type)
        return hit[1] if hit and hit[0] == self._time_id_str() else False

    def save_precomputed_features(self, feature_type, data):
        """core/models.py:310-319: one `precomputed_<type>_<time id>.pickle` per feature type in tmp/."""
        import glob
        import pickle
        self._precomputed = getattr(self, "_precomputed", {})
        self._precomputed[feature_type] = (self._time_id_str(), data)
        d = getattr(self, "tmp_out_dir", None)
        if d is None:
            return
        for old in glob.glob(os.path.join(d, "precomputed_{}*".format(feature_type))):
            os.remove(old)
        with open(os.path.join(d, "precomputed_{}_{}.pickle".format(feature_type, self._time_id_str())), "wb") as f:
            pickle.dump(data, f)

    def plot_and_send_notification_for(self, metrics_list):
        """core/models.py:259-269 draws `plots/evaluation_plots.png` and posts it to Slack; plotting and the notifier are
        outside the hot-path scope, so the same per-metric plot data goes to
        `plots/evaluation_metrics.json` instead.  Returns the file path."""
        import json
        report = {}
        for m, metric in metrics_list.items():
            data = metric.get_data_for_plot()
            report[m] = {"plot_type": getattr(metric, "plot_type", None),
                         "data": np.asarray(data).tolist() if not isinstance(data, dict) else
                                 {k: np.asarray(v).tolist() for k, v in data.items()},
                         "last_value": getattr(metric, "last_value_repr", None)}
        path = os.path.join(getattr(self, "plots_out_dir", None) or ".", "evaluation_metrics.json")
        with open(path, "w") as f:
            json.dump(report, f, indent=2, sort_keys=True, default=str)
        print("[Metrics] {}".format(", ".join("{}={}".format(m, r["last_value"]) for m, r in report.items())))
        return path

    def clean_up_tmp_dir(self):
        """core/models.py:271-278: remove the files (not the directories) under tmp/."""
        for the_file in os.listdir(self.tmp_out_dir):
            file_path = os.path.join(self.tmp_out_dir, the_file)
            try:
                if os.path.isfile(file_path):
                    os.unlink(file_path)
            except Exception as e:      # noqa: BLE001 — the reference prints and carries on
                print(e)

    # ---- TensorFlow checkpoints of the reference (core/models.py:321-344) ---------------------------
    def load_tf_checkpoint(self, prefix):
        """Weights, Adam moments and step counters from `tf.train.Checkpoint(transformer=, optimizer=)` files."""
        params, m, v, info = tf_checkpoint.import_checkpoint(prefix, self._layout_hps())
        e = self._train_engine
        e.load_params(params)
        if m and v:
            e.load_adam(m, v)
        if "optimizer_iter" in info:
            e.set_step(info["optimizer_iter"])
        self.current_step = info.get("current_step", info.get("optimizer_iter", self.current_step))

    def save_tf_checkpoint(self, prefix):
        e = self._train_engine
        like = init_params(self._layout_hps(), 0)
        m, v = e.export_adam(like)
        tf_checkpoint.export_checkpoint(prefix, self._layout_hps(), {k: t.numpy() for k, t in e.export_params(like).items()},
                                        {k: t.numpy() for k, t in m.items()}, {k: t.numpy() for k, t in v.items()},
                                        current_step=self.current_step, optimizer_iter=e.step_count())
        return prefix

    # ---- checkpoints (own format: the flat arenas + step) ------------------------------------
    def state_dict(self):
        e = self._train_engine
        return {"params": e.params.cpu(), "adam_m": e.adam_m.cpu(), "adam_v": e.adam_v.cpu(),
                "step": e.step_count(), "current_step": self.current_step,
                "table": {k: list(v) for k, v in e.table.items()},
                "metrics": e.tensor("stats_run", dtype=torch.float64).cpu().clone(),
                "hps": {k: (list(v) if isinstance(v, (list, tuple)) else v) for k, v in self.hps.items()}}

    def load_state_dict(self, sd):
        e = self._train_engine
        table = {k: tuple(int(i) for i in v) for k, v in sd["table"].items()}
        if table != e.table:
            raise ValueError("checkpoint was written for a different model layout (parameter table mismatch: "
                             f"{len(table)} tensors in the file, {len(e.table)} in this model)")
        e.params.copy_(sd["params"]); e.adam_m.copy_(sd["adam_m"]); e.adam_v.copy_(sd["adam_v"])
        e.set_step(sd["step"])
        if "metrics" in sd:
            e.tensor("stats_run", dtype=torch.float64).copy_(sd["metrics"])
        self.current_step = sd["current_step"]


models_by_name = {Transformer.name: Transformer}


def get_model_by_name(model_name):
    """models/__init__.py:16-17"""
    return models_by_name[model_name]

# models/sketchformer_b200.py — the ONE file a maintainer adds to the reference tree (leosampaio/sketchformer) to run
# its train.py / evaluate-metrics.py / experiments on the B200 path.  The registry (models/__init__.py:9-17) picks up
# every BaseModel subclass with a `name`; nothing else in the reference changes.
#
#   python train.py sketch-transformer-b200 --dataset <dir> -o <out> --id run1 \
#          --hparams d_model=256,dff=1024,dropout_rate=0.0 --base-hparams batch_size=128
#
# BaseModel keeps what it owns — train() (core/models.py:163-197), status reports, quick-metric history, the slow
# metrics and their threads, plots, Slack — and this class hands the device work to sketchformer_b200 (libskf.so).
# The weights live in the engine's HBM arenas, not in tf.Variables, so the three checkpoint hooks of BaseModel
# (core/models.py:321-358) are overridden to save / restore those arenas at the reference's cadence.
import os

from core.models import BaseModel
from models.evaluation_mixin import TransformerMetricsMixin
from sketchformer_b200.model import Transformer as _B200


class TransformerB200(BaseModel, TransformerMetricsMixin):
    """Same hparams, quick metrics and slow metrics as models/sketchformer.py:17-53; device path = libskf.so."""
    name = 'sketch-transformer-b200'
    quick_metrics = list(_B200.quick_metrics)
    slow_metrics = list(_B200.slow_metrics)

    @classmethod
    def specific_default_hparams(cls):
        from utils.hparams import HParams
        return HParams(**_B200.specific_default_hparams().values())

    # ---- construction (BaseModel.__init__ calls these two, core/models.py:96-97) ------------------
    def build_model(self):
        # out_dir=None: BaseModel has already made the directories; the engine is told where the weights go below
        self._impl = _B200(self.hps, self.dataset, None, self.experiment_id)
        self._impl.wgt_out_dir = self.wgt_out_dir
        self._impl.batches_per_epoch = self.batches_per_epoch

    def prepare_checkpointing(self):
        pass            # no tf.train.Checkpoint: see save_checkpoint_if_its_time / restore_checkpoint_if_exists

    # ---- the train step (models/sketchformer.py:351-359) --------------------------------------
    def train_on_batch(self, batch):
        return self._impl.train_on_batch(batch)

    def prepare_for_start_of_epoch(self):
        pass

    def prepare_for_end_of_epoch(self):
        self._impl.prepare_for_end_of_epoch()

    # ---- checkpoints at BaseModel's cadence (core/models.py:333-358) ------------------------------
    def save_checkpoint_if_its_time(self):
        # same rule as core/models.py:346-358 on the step train() has just counted: rolling ckpt-<step>.pt (2 kept)
        # every safety_save epochs, permanent step<step>.pt every save_every epochs, in weights/
        self._impl.current_step = int(self.current_step.numpy())
        self._impl.save_checkpoint_if_its_time()

    def restore_checkpoint_if_exists(self, checkpoint):
        if self._impl.restore_checkpoint_if_exists(checkpoint):
            self.current_step.assign(self._impl.current_step)

    # ---- inference (models/sketchformer.py:162-168, 201-311) ---------------------------------
    def encode_from_seq(self, inp_seq):
        return self._impl.encode_from_seq(inp_seq)

    def predict_class(self, inp_seq):
        return self._impl.predict_class(inp_seq)

    def predict(self, inp_seq):
        return self._impl.predict(inp_seq)          # KV-cached greedy decode on the device

    def predict_from_embedding(self, emb, expected_len=None):
        return self._impl.predict_from_embedding(emb, expected_len)

"""A torch-backed stand-in for the slice of the TensorFlow-2.1 API that the reference's layer code touches
(test infrastructure: only tests/ and the oracle/ generators import it).

TensorFlow cannot be installed in this image, so the reference's model cannot run as shipped.  Its *layer code*
(`builders/utils.py`, `builders/layers/transformer.py`, `builders/schedulers.py`) however only calls a few dozen TF
primitives with textbook semantics.  `build_modules()` makes modules named `tensorflow`, `tensorflow.keras` and
`tensorflow.keras.backend` that implement exactly those primitives on torch tensors; `load_reference_builders()` /
`load_reference_model()` put them into `sys.modules` (restored afterwards) and EXECUTE the reference's own files from /root/reference against them.  What this
pins is the wiring of the reference (argument order v/k/q, which mask goes where and with which polarity, residual /
LayerNorm order, sqrt(d_model) scaling, the transposes of split_heads and DenseExpander, the softmax axes of the
pooling layer, the schedule's formula); what it cannot pin is TensorFlow's own kernels, whose semantics are restated
here from the public API documentation:

  Dense: y = x @ kernel + bias over the last axis, kernel (in, units); activation 'relu' / 'softmax' / None
  Embedding: table[ids];  LayerNormalization(epsilon): last axis, biased variance, gamma * xhat + beta
  Dropout: identity unless training (the shim refuses training=True: TF's RNG stream is not reproducible)
  K.dot(x, W) = tensordot over x's last and W's first axis;  K.softmax / tf.nn.softmax over `axis`
  tf.linalg.band_part(x, -1, 0) = lower triangle;  tf.cast(x, tf.float32) really rounds to float32

Layers keep TF's lazy `build(input_shape)` protocol, `add_weight`, attribute tracking is not needed.  Weights are
plain tensors that the tests overwrite with the oracle's parameters after the first (building) call.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
WORK = torch.float64          # dtype of variables and of tensors whose dtype TF would infer as its default float


class _DType:
    def __init__(self, t):
        self.t = t


float32, float64, int32, int64 = _DType(torch.float32), _DType(torch.float64), _DType(torch.int32), _DType(torch.int64)
newaxis = None


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(dtype)
    return torch.as_tensor(np.asarray(x), dtype=dtype)


def cast(x, dtype):
    return _t(x).to(dtype.t if isinstance(dtype, _DType) else dtype)      # `loss_.dtype` is a torch dtype here


def shape(x):
    return tuple(_t(x).shape)


def reshape(x, shp):
    return _t(x).reshape(tuple(int(s) for s in shp))


def transpose(x, perm=None):
    return _t(x).permute(*perm) if perm is not None else _t(x).T


def matmul(a, b, transpose_b=False):
    return torch.matmul(a, b.transpose(-1, -2) if transpose_b else b)


def expand_dims(x, axis):
    return _t(x).unsqueeze(axis)


class _TFTensor(torch.Tensor):
    """torch.Tensor that, like a tf.Tensor, multiplies / adds with python lists (make_dummy_input, the SOS seed)."""

    @staticmethod
    def _lift(self, o):
        return torch.as_tensor(o, dtype=self.dtype) if isinstance(o, (list, tuple)) else o

    def __mul__(self, o):
        return torch.Tensor.__mul__(self, _TFTensor._lift(self, o))

    def __add__(self, o):
        return torch.Tensor.__add__(self, _TFTensor._lift(self, o))

    def sum(self, *args, **kwargs):        # np.sum(tensor) calls tensor.sum(axis=None, out=None) (sketchformer.py:290)
        kwargs.pop("out", None)
        axis = kwargs.pop("axis", None)
        base = self.as_subclass(torch.Tensor)
        return base.sum(*args, **kwargs) if axis is None else base.sum(axis, *args, **kwargs)


def ones(shp, dtype=None):
    return torch.ones(tuple(int(s) for s in shp), dtype=(dtype.t if dtype else torch.float32)).as_subclass(_TFTensor)


def zeros_like(x):
    return torch.zeros_like(_t(x))


def concat(values, axis=0):
    vals = [_t(v).as_subclass(torch.Tensor) for v in values]
    dt = vals[0].dtype
    for v in vals[1:]:
        dt = torch.promote_types(dt, v.dtype)
    return torch.cat([v.to(dt) for v in vals], dim=axis)


def reduce_sum(x, axis=None):
    return _t(x).sum() if axis is None else _t(x).sum(axis)


def function(*a, **k):            # @tf.function(input_signature=...): the traced function is never called here
    return (lambda f: f) if not (len(a) == 1 and callable(a[0]) and not k) else a[0]


class TensorSpec:
    def __init__(self, *a, **k):
        pass


def maximum(a, b):
    return torch.maximum(_t(a), _t(b))


def squeeze(x, axis=None):
    return _t(x).squeeze() if axis is None else _t(x).squeeze(axis)


def reduce_mean(x, axis=None):
    return _t(x).mean() if axis is None else _t(x).mean(axis)


def argmax(x, axis=-1):
    return _t(x).argmax(axis).as_subclass(_TFTensor)


def constant(x, dtype=None):
    return _t(x, dtype.t if dtype else None)


def custom_gradient(fn):          # decorator on builders/utils.py:gather, which the layer code never calls
    return fn


math = types.SimpleNamespace(
    sqrt=lambda x: torch.sqrt(_t(x)),
    rsqrt=lambda x: torch.rsqrt(_t(x, None) if isinstance(x, torch.Tensor) else torch.as_tensor(float(x), dtype=torch.float32)),
    minimum=lambda a, b: torch.minimum(_t(a), _t(b)),
    equal=lambda a, b: _t(a) == b,
    logical_not=lambda a: ~_t(a),
    pow=lambda a, b: torch.pow(_t(a), b),
    floor=lambda a: torch.floor(_t(a)),
)
linalg = types.SimpleNamespace(band_part=lambda x, lo, hi: (torch.tril(x) if (lo, hi) == (-1, 0) else
                                                            (_ for _ in ()).throw(NotImplementedError("band_part"))))
def _sparse_ce_from_logits(labels, logits):
    lsm = torch.log_softmax(_t(logits), dim=-1)
    return -lsm.gather(-1, _t(labels).long().unsqueeze(-1)).squeeze(-1)


nn = types.SimpleNamespace(softmax=lambda x, axis=-1: torch.softmax(_t(x), dim=axis),
                           sparse_softmax_cross_entropy_with_logits=lambda labels, logits: _sparse_ce_from_logits(labels, logits))
# tf.losses.mean_squared_error == tf.keras.losses.MSE: mean of the squared difference over the LAST axis
losses = types.SimpleNamespace(mean_squared_error=lambda a, b: ((_t(b) - _t(a)) ** 2).mean(-1))


class SparseCategoricalCrossentropy:
    """tf.keras.losses.SparseCategoricalCrossentropy.  from_logits=True: -log_softmax(pred)[y].  from_logits=False:
    Keras clips the probabilities to [1e-7, 1 - 1e-7] and takes -log p[y] (eager semantics; in graph mode TF-2.1
    back-tracks a Softmax op to its logits, SURVEY A.3 - the two agree unless p < 1e-7).  reduction 'none' keeps the
    per-element values, the default averages over all of them."""

    def __init__(self, from_logits=False, reduction="auto", name=None):
        self.from_logits, self.reduction = from_logits, reduction

    def __call__(self, y_true, y_pred):
        y = _t(y_true).long()
        if y.dim() == _t(y_pred).dim():          # labels (B, 1) against predictions (B, C)
            y = y.squeeze(-1)
        if self.from_logits:
            v = _sparse_ce_from_logits(y, y_pred)
        else:
            pr = torch.clamp(_t(y_pred), 1e-7, 1 - 1e-7)
            v = -torch.log(pr.gather(-1, y.unsqueeze(-1)).squeeze(-1))
        return v if self.reduction == "none" else v.mean()


class _Mean:
    def __init__(self, name=None):
        self.total, self.count = 0.0, 0

    def __call__(self, v):
        self.total += float(_t(v).double().mean()); self.count += 1

    def result(self):
        return torch.tensor(self.total / max(1, self.count))

    def reset_states(self):
        self.total, self.count = 0.0, 0


class _SparseCategoricalAccuracy(_Mean):
    def __call__(self, y_true, y_pred):
        y = _t(y_true).long()
        if y.dim() == _t(y_pred).dim():
            y = y.squeeze(-1)
        hit = (_t(y_pred).argmax(-1) == y).double()
        self.total += float(hit.sum()); self.count += hit.numel()


class GradientTape:
    """tf.GradientTape as used by model_trainer (models/sketchformer.py:331-347): gradients of a scalar with respect
    to a list of variables - here torch leaf tensors with requires_grad."""

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def gradient(self, target, sources):
        return torch.autograd.grad(target, list(sources), allow_unused=True)


class _RecordedOptimizer:
    """tf.keras.optimizers.Adam: records the constructor arguments build_model() passes, and implements
    apply_gradients with the update rule TensorFlow documents for Adam (training_ops ApplyAdam):
        t = iterations + 1;  lr_t = lr(iterations) * sqrt(1 - beta_2^t) / (1 - beta_1^t)
        m = beta_1 m + (1 - beta_1) g;  v = beta_2 v + (1 - beta_2) g^2;  var -= lr_t * m / (sqrt(v) + epsilon)
    (the schedule is evaluated at `iterations` BEFORE the increment).  This restates TensorFlow, not the reference."""

    def __init__(self, *args, **kwargs):
        self.args, self.kwargs = args, kwargs
        self.iterations, self.slots = 0, {}

    def apply_gradients(self, grads_and_vars):
        lr = self.args[0]
        b1, b2, eps = self.kwargs["beta_1"], self.kwargs["beta_2"], self.kwargs["epsilon"]
        lr = float(lr(torch.tensor(float(self.iterations), dtype=torch.float32))) if callable(lr) else float(lr)
        t = self.iterations + 1
        lr_t = lr * (1 - b2 ** t) ** 0.5 / (1 - b1 ** t)
        with torch.no_grad():
            for g, var in grads_and_vars:
                if g is None:
                    continue
                m, v = self.slots.setdefault(id(var), (torch.zeros_like(var), torch.zeros_like(var)))
                m.mul_(b1).add_(g, alpha=1 - b1)
                v.mul_(b2).addcmul_(g, g, value=1 - b2)
                var.sub_(lr_t * m / (v.sqrt() + eps))
        self.iterations = t


# ---- tf.keras -------------------------------------------------------------------------------------------------------
class Layer:
    def __init__(self, **kwargs):
        self.built = False

    def add_weight(self, name=None, shape=None, initializer=None, trainable=True):
        return torch.zeros(tuple(int(s) for s in shape), dtype=WORK)

    def build(self, input_shape):
        self.built = True

    def __call__(self, *args, **kwargs):
        if not getattr(self, "built", False):
            first = args[0] if args else next(iter(kwargs.values()))
            self.build(tuple(first.shape))
            self.built = True
        return self.call(*args, **kwargs)


class Dense(Layer):
    def __init__(self, units, activation=None, **kwargs):
        super().__init__(**kwargs)
        self.units, self.activation = int(units), activation
        self.kernel = self.bias = None

    def build(self, input_shape):
        self.kernel = torch.zeros(int(input_shape[-1]), self.units, dtype=WORK)
        self.bias = torch.zeros(self.units, dtype=WORK)
        self.built = True

    def call(self, x):
        y = torch.matmul(x.to(self.kernel.dtype), self.kernel) + self.bias
        if self.activation == "relu":
            y = torch.relu(y)
        elif self.activation == "softmax":
            y = torch.softmax(y, dim=-1)
        elif self.activation is not None:
            raise NotImplementedError(self.activation)
        return y


class Embedding(Layer):
    def __init__(self, input_dim, output_dim, **kwargs):
        super().__init__(**kwargs)
        self.embeddings = torch.zeros(int(input_dim), int(output_dim), dtype=WORK)
        self.built = True

    def call(self, ids):
        return self.embeddings[ids.long()]


class LayerNormalization(Layer):
    def __init__(self, epsilon=1e-3, **kwargs):
        super().__init__(**kwargs)
        self.epsilon = epsilon
        self.gamma = self.beta = None

    def build(self, input_shape):
        self.gamma = torch.ones(int(input_shape[-1]), dtype=WORK)
        self.beta = torch.zeros(int(input_shape[-1]), dtype=WORK)
        self.built = True

    def call(self, x):
        mu = x.mean(-1, keepdim=True)
        var = ((x - mu) ** 2).mean(-1, keepdim=True)
        return (x - mu) * torch.rsqrt(var + self.epsilon) * self.gamma + self.beta


class Dropout(Layer):
    def __init__(self, rate, **kwargs):
        super().__init__(**kwargs)
        self.rate = rate
        self.built = True

    def call(self, x, training=None):
        if training and self.rate > 0:
            raise NotImplementedError("the shim runs inference graphs only (TF's dropout RNG is not reproducible)")
        return x


class Sequential(Layer):
    def __init__(self, layers=None, **kwargs):
        super().__init__(**kwargs)
        self.layers = list(layers or [])
        self.built = True

    def call(self, x):
        for lay in self.layers:
            x = lay(x)
        return x


class LearningRateSchedule:
    def __init__(self, *a, **k):
        pass


backend = types.SimpleNamespace(
    dot=lambda x, w: torch.matmul(x.to(w.dtype), w),
    bias_add=lambda x, b: x + b,
    tanh=torch.tanh,
    softmax=lambda x, axis=-1: torch.softmax(x, dim=axis),
    sum=lambda x, axis=None, keepdims=False: torch.sum(x, dim=axis, keepdim=keepdims),
)


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def build_modules():
    layers = _module("tensorflow.keras.layers", Layer=Layer, Dense=Dense, Embedding=Embedding,
                     LayerNormalization=LayerNormalization, Dropout=Dropout)
    schedules = types.SimpleNamespace(LearningRateSchedule=LearningRateSchedule)
    kback = _module("tensorflow.keras.backend", **backend.__dict__)
    keras = _module("tensorflow.keras", layers=layers, Sequential=Sequential, backend=kback,
                    optimizers=types.SimpleNamespace(schedules=schedules, Adam=_RecordedOptimizer, SGD=_RecordedOptimizer),
                    losses=types.SimpleNamespace(SparseCategoricalCrossentropy=SparseCategoricalCrossentropy,
                                                 MSE=losses.mean_squared_error, MAE=None),
                    metrics=types.SimpleNamespace(Mean=_Mean, SparseCategoricalAccuracy=_SparseCategoricalAccuracy))
    tf = _module("tensorflow", keras=keras, float32=float32, float64=float64, int32=int32, int64=int64, newaxis=newaxis,
                 cast=cast, shape=shape, reshape=reshape, transpose=transpose, matmul=matmul, expand_dims=expand_dims,
                 ones=ones, maximum=maximum, squeeze=squeeze, reduce_mean=reduce_mean, argmax=argmax,
                 constant=constant, custom_gradient=custom_gradient, math=math, linalg=linalg, nn=nn, losses=losses,
                 zeros_like=zeros_like, concat=concat, reduce_sum=reduce_sum, function=function, TensorSpec=TensorSpec,
                 GradientTape=GradientTape)
    return {"tensorflow": tf, "tensorflow.keras": keras, "tensorflow.keras.backend": kback,
            "tensorflow.keras.layers": layers}


def _exec(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference_builders():
    """Executes builders/utils.py, builders/schedulers.py and builders/layers/transformer.py of the reference on the
    shim; returns (utils, schedulers, transformer) modules.  sys.modules is restored afterwards."""
    names = list(build_modules()) + ["builders", "builders.utils", "builders.schedulers", "builders.layers",
                                     "builders.layers.transformer"]
    saved = {k: sys.modules.get(k) for k in names}
    try:
        sys.modules.update(build_modules())
        pkg = types.ModuleType("builders"); pkg.__path__ = [os.path.join(REF, "builders")]
        sub = types.ModuleType("builders.layers"); sub.__path__ = [os.path.join(REF, "builders", "layers")]
        sys.modules["builders"], sys.modules["builders.layers"] = pkg, sub
        utils = _exec("builders.utils", os.path.join(REF, "builders", "utils.py"))
        sched = _exec("builders.schedulers", os.path.join(REF, "builders", "schedulers.py"))
        tr = _exec("builders.layers.transformer", os.path.join(REF, "builders", "layers", "transformer.py"))
        return utils, sched, tr
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def load_reference_model():
    """Executes models/sketchformer.py (class Transformer) on the shim together with builders/{losses,
    keras_metrics,schedulers,utils,layers/transformer}.py.  `core.models.BaseModel`, `utils.hparams.HParams` and the
    metrics mixin - training loop, checkpoints, plots - are empty stand-ins: the methods under test (build_model,
    call, encode, decode, classify_from_embedding, encode_from_seq, predict*, make_dummy_input) do not use them.
    Returns (Transformer class, builders package)."""
    names = list(build_modules()) + ["builders", "builders.utils", "builders.schedulers", "builders.layers",
                                     "builders.layers.transformer", "builders.losses", "builders.keras_metrics", "utils",
                                     "utils.hparams", "core", "core.models", "models", "models.evaluation_mixin",
                                     "models.sketchformer"]
    saved = {k: sys.modules.get(k) for k in names}
    try:
        sys.modules.update(build_modules())
        pkg = types.ModuleType("builders"); pkg.__path__ = [os.path.join(REF, "builders")]
        sub = types.ModuleType("builders.layers"); sub.__path__ = [os.path.join(REF, "builders", "layers")]
        sys.modules["builders"], sys.modules["builders.layers"] = pkg, sub
        pkg.utils = _exec("builders.utils", os.path.join(REF, "builders", "utils.py"))
        pkg.schedulers = _exec("builders.schedulers", os.path.join(REF, "builders", "schedulers.py"))
        pkg.losses = _exec("builders.losses", os.path.join(REF, "builders", "losses.py"))
        pkg.keras_metrics = _exec("builders.keras_metrics", os.path.join(REF, "builders", "keras_metrics.py"))
        sub.transformer = _exec("builders.layers.transformer", os.path.join(REF, "builders", "layers", "transformer.py"))
        pkg.layers = sub
        sys.modules["utils"] = _module("utils"); sys.modules["utils"].__path__ = []
        sys.modules["utils.hparams"] = _module("utils.hparams", HParams=dict)
        sys.modules["core"] = _module("core"); sys.modules["core"].__path__ = []
        sys.modules["core.models"] = _module("core.models", BaseModel=type("BaseModel", (), {}))
        models = _module("models"); models.__path__ = [os.path.join(REF, "models")]
        sys.modules["models"] = models
        sys.modules["models.evaluation_mixin"] = _module("models.evaluation_mixin",
                                                         TransformerMetricsMixin=type("TransformerMetricsMixin", (), {}))
        mod = _exec("models.sketchformer", os.path.join(REF, "models", "sketchformer.py"))
        return mod.Transformer, pkg
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def available():
    return os.path.exists(os.path.join(REF, "builders", "layers", "transformer.py"))

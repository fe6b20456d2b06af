"""Run the reference's OWN floating-point code -- `/root/reference/KGEkeras/models.py` and the loss functions of
`utils.py` -- UNMODIFIED, over a minimal `tensorflow` namespace backed by torch CPU tensors (TEST INFRASTRUCTURE).

Why: TensorFlow, rdflib and spaCy are not installable here, so the reference cannot be imported as it stands and the
floating-point half of the oracle (`kge_oracle.score`, the losses, the gradients) used to be a restatement checked only
against other restatements.  This module closes that gap as far as the image allows:

  * the two reference files are loaded WHERE THEY LIE with `importlib` (nothing is copied), while `sys.modules` holds
    stand-ins for `tensorflow[.keras.*]`, `rdflib[.namespace]` and `spacy` for the duration of the import;
  * every `tf.*` / `K.*` call the hot path makes (models.py:134-162 `call`, 171-237 and 383-489 the `func` bodies,
    utils.py:288-319 the losses) is mapped one-to-one onto the torch op of the same meaning, so the OP SEQUENCE, the
    operand order, the branches (`gamma > 0`), the constants (`EPSILON+1`, `embedding_range / pi`) and the broadcasting
    are the reference's, executed from the reference's source;
  * gradients come from torch autograd through that op sequence -- the role TensorFlow's autodiff plays for
    `model.add_loss(loss)` (tests.py:79-83).

What this pins and what it cannot: formulas, operand order, shapes, branches and derivative structure are pinned by the
reference source.  The ROUNDING of each primitive is torch's CPU kernel, not TensorFlow's Eigen kernel: elementwise
fp32 ops are IEEE and agree bit for bit, reductions (`reduce_sum`, `norm`, `fft`) may associate differently, which is
why the fixtures made from this module (`tests/golden/make_golden_fp.py`) are compared inside the `1e-5 * sum|terms|`
bar and carry an fp64 run of the same source as the truth.  Third-party pieces the reference only CALLS are restated
here from their published definitions and say so: Keras `binary_crossentropy` (tf.keras 2.x backend: clip to
[1e-7, 1-1e-7], `-(y log(p+eps) + (1-y) log(1-p+eps))`, mean over the last axis), `GlorotUniform`, `Dense`, `Embedding`,
and Keras' default `predict` batch of 32 rows.  Dropout is the identity here (its masks come from TensorFlow's RNG stream).

Only tests/ and tests/golden/make_golden_fp.py import this file; `/root/reference` does not exist on the GPU box, so
callers check `available()` and the GPU tests read the committed fixture instead.
"""
from __future__ import annotations

import contextlib
import importlib.util
import os
import sys
import types

import numpy as np

REF_DIR = "/root/reference/KGEkeras"
REF_MODELS = os.path.join(REF_DIR, "models.py")
REF_UTILS = os.path.join(REF_DIR, "utils.py")


def available() -> bool:
    return os.path.isfile(REF_MODELS) and os.path.isfile(REF_UTILS)


# ----------------------------------------------------------------------------------------------------------------
# the stand-in namespace
# ----------------------------------------------------------------------------------------------------------------
class _State:
    real = None      # torch.float32 | torch.float64, set by load()
    cplx = None
    rng = None       # numpy Generator behind GlorotUniform


def _torch():
    import torch
    return torch


def _as_tensor(x):
    torch = _torch()
    if isinstance(x, torch.Tensor):
        return x
    a = np.asarray(x)
    if a.dtype.kind == "f":
        return torch.as_tensor(a.astype(np.float64)).to(_State.real)
    return torch.as_tensor(a)


class _ComplexTag:
    """`tf.complex64`: resolves to complex64 in the fp32 run and complex128 in the fp64 truth run."""


def _build_tf():
    torch = _torch()
    tf = types.ModuleType("tensorflow")

    def norm(tensor, ord="euclidean", axis=None, keepdims=None, name=None):
        # tf.norm: 'euclidean'/2 -> sqrt(reduce_sum(x * conj(x))), 1 -> reduce_sum(abs(x)),
        # inf -> reduce_max(abs(x)), p -> pow(reduce_sum(pow(abs(x), p)), 1/p); axis=None reduces over ALL elements
        t = _as_tensor(tensor)
        red = (lambda v: torch.sum(v)) if axis is None else (lambda v: torch.sum(v, dim=axis))
        if ord in ("euclidean", 2, 2.0):
            return torch.sqrt(red(t * t))
        if ord in (1, 1.0):
            return red(torch.abs(t))
        if ord == np.inf:
            return torch.amax(torch.abs(t)) if axis is None else torch.amax(torch.abs(t), dim=axis)
        return torch.pow(red(torch.pow(torch.abs(t), ord)), 1.0 / ord)

    def reduce_sum(x, axis=None, keepdims=False):
        x = _as_tensor(x)
        return torch.sum(x) if axis is None else torch.sum(x, dim=axis, keepdim=keepdims)

    def reduce_mean(x, axis=None, keepdims=False):
        x = _as_tensor(x)
        return torch.mean(x) if axis is None else torch.mean(x, dim=axis, keepdim=keepdims)

    def split(value, num_or_size_splits, axis=0):
        value = _as_tensor(value)
        if isinstance(num_or_size_splits, int):
            assert value.shape[axis] % num_or_size_splits == 0   # tf.split raises on an uneven split
            return list(torch.split(value, value.shape[axis] // num_or_size_splits, dim=axis))
        return list(torch.split(value, list(num_or_size_splits), dim=axis))

    def cast(x, dtype):
        x = _as_tensor(x)
        if dtype is _ComplexTag:
            return x.to(_State.cplx)
        return x.to(dtype)

    tf.norm = norm
    tf.reduce_sum = reduce_sum
    tf.reduce_mean = reduce_mean
    tf.split = split
    tf.cast = cast
    tf.complex64 = _ComplexTag
    tf.unstack = lambda x, axis=0: list(torch.unbind(_as_tensor(x), dim=axis))
    tf.concat = lambda values, axis: torch.cat([_as_tensor(v) for v in values], dim=axis)
    tf.multiply = lambda a, b: torch.mul(a, b)
    tf.abs = lambda x: torch.abs(_as_tensor(x))
    tf.square = lambda x: torch.square(_as_tensor(x))
    tf.where = lambda c, a, b: torch.where(c, a, b)
    tf.reshape = lambda x, shape: torch.reshape(_as_tensor(x), tuple(shape))

    tf.math = types.ModuleType("tensorflow.math")
    tf.math.real = lambda x: torch.real(x)
    tf.math.conj = lambda x: torch.conj(x)
    tf.math.sin = lambda x: torch.sin(_as_tensor(x))
    tf.math.cos = lambda x: torch.cos(_as_tensor(x))
    tf.math.atan2 = lambda y, x: torch.atan2(y, x)
    tf.math.log = lambda x: torch.log(_as_tensor(x))
    tf.math.exp = lambda x: torch.exp(_as_tensor(x))

    tf.signal = types.ModuleType("tensorflow.signal")
    tf.signal.fft = lambda x: torch.fft.fft(x, dim=-1)      # tf.signal.fft: over the innermost dimension
    tf.signal.ifft = lambda x: torch.fft.ifft(x, dim=-1)

    tf.nn = types.ModuleType("tensorflow.nn")
    tf.nn.relu = lambda x: torch.relu(_as_tensor(x))

    # ---- keras ------------------------------------------------------------------------------------------------
    keras = types.ModuleType("tensorflow.keras")

    class GlorotUniform:
        """Keras GlorotUniform: U(-l, l), l = sqrt(6 / (fan_in + fan_out)).  TensorFlow's random stream is not
        reproducible here; values come from a seeded NumPy generator (fixtures store the tables they used)."""

        def __call__(self, shape):
            lim = np.sqrt(6.0 / (shape[0] + shape[1]))
            return _as_tensor(_State.rng.uniform(-lim, lim, size=shape).astype(np.float32))

    class _Layer:
        def __init__(self, *a, **k):
            self.trainable = True

    class Embedding(_Layer):
        def __init__(self, input_dim, output_dim, embeddings_initializer=None, embeddings_regularizer=None,
                     weights=None, name=None, **kw):
            super().__init__()
            if weights is not None:
                w = _as_tensor(np.asarray(weights[0]))
                assert tuple(w.shape) == (input_dim, output_dim), (tuple(w.shape), input_dim, output_dim)
            else:
                w = (embeddings_initializer or GlorotUniform())((input_dim, output_dim))
            self.embeddings = w.clone().detach().requires_grad_(True)
            self.embeddings_regularizer = embeddings_regularizer
            self.name = name

        def __call__(self, ids):
            return self.embeddings[_as_tensor(ids).long()]

    class Dropout(_Layer):
        def __init__(self, rate, **kw):
            super().__init__()
            self.rate = rate

        def __call__(self, x, training=None):
            return x      # inference semantics; see the module docstring

    _ACT = {None: lambda z: z, "linear": lambda z: z, "tanh": torch.tanh, "relu": torch.relu,
            "sigmoid": torch.sigmoid, "softplus": torch.nn.functional.softplus,
            "softsign": torch.nn.functional.softsign, "elu": torch.nn.functional.elu,
            "selu": torch.nn.functional.selu, "exponential": torch.exp,
            "softmax": lambda z: torch.softmax(z, dim=-1)}

    class Dense(_Layer):
        """Keras Dense: activation(x @ kernel + bias), kernel (in, units) built at the first call."""

        def __init__(self, units, activation=None, **kw):
            super().__init__()
            self.units, self.activation = units, activation
            self.kernel = self.bias = None

        def build(self, fan_in):
            self.kernel = GlorotUniform()((fan_in, self.units)).clone().detach().requires_grad_(True)
            self.bias = torch.zeros(self.units, dtype=_State.real, requires_grad=True)

        def set_weights(self, kernel, bias):
            self.kernel = _as_tensor(kernel).clone().detach().requires_grad_(True)
            self.bias = _as_tensor(bias).clone().detach().requires_grad_(True)

        def __call__(self, x):
            if self.kernel is None:
                self.build(x.shape[-1])
            act = self.activation if callable(self.activation) else _ACT[self.activation]
            return act(x @ self.kernel + self.bias)

    class _Unserved(_Layer):
        """Layers only the out-of-scope convolutional models construct."""

    class Model:
        def __init__(self, *a, **k):
            pass

        def __call__(self, inputs, training=False):
            return self.call(_as_tensor(inputs), training=training)

        def predict(self, x, batch_size=32, verbose=0):
            """Keras Model.predict: batches of `batch_size` rows (default 32), outputs concatenated."""
            x = np.asarray(x)
            with torch.no_grad():
                outs = [np.atleast_1d(self(x[i:i + batch_size]).numpy()) for i in range(0, len(x), batch_size)]
            return np.concatenate(outs) if outs else np.zeros((0,), np.float32)

    layers = types.ModuleType("tensorflow.keras.layers")
    layers.Layer, layers.Embedding, layers.Dense, layers.Dropout = _Layer, Embedding, Dense, Dropout
    for nm in ("Lambda", "Multiply", "Reshape", "Concatenate", "BatchNormalization", "Conv2D", "Activation",
               "Conv3D", "Flatten", "Dot"):
        setattr(layers, nm, type(nm, (_Unserved,), {}))
    models = types.ModuleType("tensorflow.keras.models")
    models.Model = Model
    initializers = types.ModuleType("tensorflow.keras.initializers")
    initializers.GlorotUniform = GlorotUniform
    backend = types.ModuleType("tensorflow.keras.backend")
    backend.abs = lambda x: torch.abs(x)
    backend.clip = lambda x, min_value, max_value: torch.clamp(x, min=float(min_value), max=float(max_value))
    callbacks = types.ModuleType("tensorflow.keras.callbacks")
    callbacks.Callback = type("Callback", (), {"__init__": lambda self, *a, **k: None})
    losses = types.ModuleType("tensorflow.keras.losses")

    def binary_crossentropy(y_true, y_pred, from_logits=False, label_smoothing=0.0, axis=-1):
        # tf.keras 2.x: losses.binary_crossentropy -> K.mean(K.binary_crossentropy(y_true, y_pred), axis=-1);
        # K.binary_crossentropy (probabilities): output = clip(output, eps, 1-eps);
        # bce = y*log(output+eps) + (1-y)*log(1-output+eps); return -bce         (eps = K.epsilon() = 1e-7)
        assert not from_logits and label_smoothing == 0.0
        y_pred = _as_tensor(y_pred)
        y = torch.as_tensor(y_true, dtype=y_pred.dtype)
        eps = torch.as_tensor(1e-7, dtype=y_pred.dtype)
        out = torch.clamp(y_pred, min=eps, max=1 - eps)
        bce = y * torch.log(out + eps) + (1 - y) * torch.log(1 - out + eps)
        return torch.mean(-bce, dim=axis)

    losses.binary_crossentropy = binary_crossentropy
    keras.Model, keras.layers, keras.models, keras.initializers = Model, layers, models, initializers
    keras.backend, keras.callbacks, keras.losses = backend, callbacks, losses
    tf.keras = keras

    mods = {"tensorflow": tf, "tensorflow.keras": keras, "tensorflow.keras.models": models,
            "tensorflow.keras.layers": layers, "tensorflow.keras.backend": backend,
            "tensorflow.keras.callbacks": callbacks, "tensorflow.keras.losses": losses,
            "tensorflow.keras.initializers": initializers, "tensorflow.math": tf.math,
            "tensorflow.signal": tf.signal, "tensorflow.nn": tf.nn}
    # utils.py:14-20 imports these at module level for LiteralConverter (out of scope); only the names must exist
    rdflib = types.ModuleType("rdflib")
    for nm in ("Graph", "URIRef", "Literal"):
        setattr(rdflib, nm, type(nm, (), {}))
    rdflib.Namespace = lambda s: s
    ns = types.ModuleType("rdflib.namespace")
    ns.XSD = ns.RDF = None
    rdflib.namespace = ns
    mods.update({"rdflib": rdflib, "rdflib.namespace": ns, "spacy": types.ModuleType("spacy")})
    return mods


@contextlib.contextmanager
def _stand_ins(mods):
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    try:
        yield
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


class _NumpyWithInf:
    """models.py:394 spells `np.Inf`, removed in NumPy 2; the reference's pinned NumPy 1.x has it."""
    Inf = np.inf

    def __getattr__(self, name):
        return getattr(np, name)


def _load_file(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load(precision: int = 32, seed: int = 0):
    """Execute the two reference files and return (models_module, utils_module, tf_stand_in).

    precision 32: tensors are fp32 (the reference's dtype); 64: the same source in fp64 (the truth the bars refer to).
    The modules are private instances (never registered in sys.modules, never the repo's `KGEkeras`)."""
    torch = _torch()
    assert available(), "the reference is only mounted in the build container"
    assert precision in (32, 64)
    _State.real = torch.float32 if precision == 32 else torch.float64
    _State.cplx = torch.complex64 if precision == 32 else torch.complex128
    _State.rng = np.random.default_rng(seed)
    mods = _build_tf()
    with _stand_ins(mods):
        rm = _load_file(REF_MODELS, "_kgeb_reference_models_p%d" % precision)
        ru = _load_file(REF_UTILS, "_kgeb_reference_utils_p%d" % precision)
    rm.np = _NumpyWithInf()
    return rm, ru, mods["tensorflow"]


def set_precision(precision: int):
    """Switch the dtype new tensors are made in (the loaded modules share the stand-in namespace)."""
    torch = _torch()
    _State.real = torch.float32 if precision == 32 else torch.float64
    _State.cplx = torch.complex64 if precision == 32 else torch.complex128


def model_regularization_loss(model):
    """Keras adds `embeddings_regularizer(embeddings)` of every layer to the model's losses (models.py:80-92)."""
    total = None
    for layer in (model.entity_embedding, model.relational_embedding):
        reg = getattr(layer, "embeddings_regularizer", None)
        if reg is not None:
            term = reg(layer.embeddings)
            total = term if total is None else total + term
    return total

"""NumPy stand-in for the handful of TensorFlow / Keras primitives the reference's FACT model calls -- TEST
INFRASTRUCTURE ONLY (used by tests/golden/make_reference_golden.py in the build container).

Purpose: TensorFlow cannot be installed here, but the reference model is plain Python, so its OWN code
(/root/reference/mint/core/{fact_model,base_models,base_model_util}.py) can be executed with `import tensorflow`
resolving to this module.  That pins the COMPOSITION (layer order, einsum strings, Rearrange patterns, the
dim**-0.5 scale, concat order, residual wiring, which row the AR loop keeps) to the reference's code; the primitives
below carry their documented Keras semantics in float64:
  Dense: x @ kernel + bias, kernel [in, units] glorot-uniform, bias zeros, optional activation / kernel_initializer
  LayerNormalization(epsilon): last axis, biased variance, gamma ones / beta zeros
  Sequential, Model/Layer (__call__ -> call), add_weight, einsum, nn.softmax, tanh, pow, concat, reduce_mean, square
  pad, maximum, random.uniform, Tensor.set_shape (checked) for mint/utils/inputs_util.py (tests/golden/make_inputs_golden.py)
Nothing here is used by the product or shipped to the GPU box.
"""
import math
import types

import numpy as np

float32 = np.float32
float64 = np.float64
int32 = np.int32

_RNG = np.random.default_rng(0)
DTYPE = np.float64


def set_seed(seed):
    global _RNG
    _RNG = np.random.default_rng(seed)


class _Shape(tuple):
    def as_list(self):
        return list(self)


class Tensor(np.ndarray):
    """ndarray whose .shape also answers .as_list() (base_model_util.get_shape_list)."""

    @property
    def shape(self):
        return _Shape(np.ndarray.shape.__get__(self))


    def set_shape(self, shape):  # static-shape annotation in TF; here: check it (inputs_util.py:88-103)
        want = tuple(shape)
        got = tuple(np.ndarray.shape.__get__(self))
        if len(want) != len(got) or any(w is not None and int(w) != g for w, g in zip(want, got)):
            raise ValueError(f"set_shape{want} on a tensor of shape {got}")


class Variable(Tensor):  # einops probes tf.Tensor / tf.Variable when a module named tensorflow is loaded
    pass


def _t(x):
    return np.asarray(x, dtype=DTYPE).view(Tensor)


def convert_to_tensor(x, dtype=None):
    return _t(x)


constant = convert_to_tensor


def einsum(eq, *ops):
    return _t(np.einsum(eq, *[np.asarray(o) for o in ops]))


def concat(values, axis):
    return _t(np.concatenate([np.asarray(v) for v in values], axis=axis))


def tanh(x):
    return _t(np.tanh(np.asarray(x)))


def pow(x, y):  # noqa: A001
    return _t(np.power(np.asarray(x), y))


def square(x):
    return _t(np.square(np.asarray(x)))


def reduce_mean(x, axis=None):
    return _t(np.mean(np.asarray(x), axis=axis))


def shape(x):
    return np.array(np.asarray(x).shape)


def _softmax(x, axis=-1):
    x = np.asarray(x)
    e = np.exp(x - x.max(axis=axis, keepdims=True))
    return _t(e / e.sum(axis=axis, keepdims=True))


# ---- the primitives of mint/utils/inputs_util.py:59-107 (fact_preprocessing), dtype-preserving
def pad(x, paddings):
    a = np.asarray(x)
    return np.pad(a, [[int(p[0]), int(p[1])] for p in paddings]).view(Tensor)


def maximum(a, b):
    return np.maximum(a, b)


def _random_uniform(shape, minval=0, maxval=None, dtype=None):
    """tf.random.uniform for the one call the reference makes: a scalar int32 in [minval, maxval).  `FORCED_START`
    (set by the golden generator) replaces the draw so that both implementations slice the same window."""
    if FORCED_START is not None:
        return np.int32(FORCED_START)
    return np.int32(_RNG.integers(int(minval), int(maxval)))


FORCED_START = None
random = types.SimpleNamespace(uniform=_random_uniform)

nn = types.SimpleNamespace(softmax=_softmax, relu=lambda x: _t(np.maximum(np.asarray(x), 0)))


class _TruncatedNormal:
    def __init__(self, mean=0.0, stddev=0.05, seed=None):
        self.mean, self.stddev = mean, stddev

    def __call__(self, shp, dtype=None):
        a = _RNG.standard_normal(tuple(shp))
        bad = np.abs(a) > 2.0
        while bad.any():
            a[bad] = _RNG.standard_normal(int(bad.sum()))
            bad = np.abs(a) > 2.0
        return _t(self.mean + self.stddev * a)


def _glorot_uniform(shp):
    lim = math.sqrt(6.0 / (shp[0] + shp[1]))
    return _t(_RNG.uniform(-lim, lim, tuple(shp)))


class Layer:
    def __init__(self, *args, **kwargs):
        self._weights = {}

    def __call__(self, *args, **kwargs):
        kwargs.pop("training", None)
        return self.call(*args, **kwargs)

    def add_weight(self, name=None, shape=None, initializer=None, dtype=None, **kwargs):  # noqa: A002
        w = initializer(shape) if callable(initializer) else _t(np.zeros(tuple(shape)))
        if not hasattr(self, "_weights"):
            self._weights = {}
        self._weights[name] = w
        return w


class Dense(Layer):
    def __init__(self, units, activation=None, use_bias=True, kernel_initializer=None, **kwargs):
        super().__init__()
        self.units, self.activation, self.use_bias = units, activation, use_bias
        self.kernel_initializer = kernel_initializer
        self.kernel = None
        self.bias = None

    def call(self, x):
        x = np.asarray(x)
        if self.kernel is None:
            shp = (x.shape[-1], self.units)
            self.kernel = self.kernel_initializer(shp) if callable(self.kernel_initializer) else _glorot_uniform(shp)
            self.bias = _t(np.zeros(self.units)) if self.use_bias else None
        y = x @ np.asarray(self.kernel)
        if self.use_bias:
            y = y + np.asarray(self.bias)
        y = _t(y)
        return self.activation(y) if self.activation is not None else y


class LayerNormalization(Layer):
    def __init__(self, axis=-1, epsilon=1e-3, **kwargs):
        super().__init__()
        assert axis == -1
        self.epsilon = epsilon
        self.gamma = None
        self.beta = None

    def call(self, x):
        x = np.asarray(x)
        if self.gamma is None:
            self.gamma = _t(np.ones(x.shape[-1]))
            self.beta = _t(np.zeros(x.shape[-1]))
        mu = x.mean(-1, keepdims=True)
        var = ((x - mu) ** 2).mean(-1, keepdims=True)
        return _t((x - mu) / np.sqrt(var + self.epsilon) * np.asarray(self.gamma) + np.asarray(self.beta))


class Sequential(Layer):
    def __init__(self, layers=None, **kwargs):
        super().__init__()
        self.layers = list(layers or [])

    def call(self, x):
        for layer in self.layers:
            x = layer(x)
        return x


keras = types.SimpleNamespace(
    Model=Layer,
    Sequential=Sequential,
    layers=types.SimpleNamespace(Layer=Layer, Dense=Dense, LayerNormalization=LayerNormalization),
    initializers=types.SimpleNamespace(TruncatedNormal=_TruncatedNormal),
    activations=types.SimpleNamespace(relu=nn.relu),
    metrics=types.SimpleNamespace(Metric=Layer),   # mint/core/metrics.py subclasses it at import time (never used)
)

"""A stand-in `tensorflow` module, just large enough to EXECUTE the reference's forward-only optimizers
(Control_Toolkit/Optimizers/optimizer_random_action_tf.py, optimizer_cem_tf.py) in this container, where TensorFlow is
not installed.

TEST INFRASTRUCTURE ONLY (used by oracle/gen_golden_plan.py; never imported by the product, the GPU tests or bench.py).
The optimizers' own Python -- control flow, call order, slicing, the distribution update and shift -- runs unmodified;
only the ~15 `tf.*` ops it calls are supplied here, on torch CPU float32 tensors, following the documented TF semantics:
  tf.argsort            ascending; equal keys keep index order (TF's CPU kernel sorts with a stable top-k)
  tf.math.reduce_std    population standard deviation: sqrt(mean(|x - mean(x)|^2)) (tf.math.reduce_variance)
  tf.clip_by_value, tf.tile, tf.gather, tf.concat, tf.reduce_mean, tf.squeeze, tf.multiply, tf.zeros, tf.ones,
  tf.convert_to_tensor, tf.constant, tf.ensure_shape: their numpy namesakes.
  optimizer_cem_gmm_tf additionally: tf.norm, tf.transpose, tf.argmin (first of equal minima), tf.cast, tf.shape, tf.stack,
  and tensorflow_probability's Normal / Categorical / MixtureSameFamily (sample, mean, stddev) with injected draws.
  optimizer_rpgd_tf additionally: tf.Variable (a torch leaf tensor with .assign), tf.GradientTape (torch autograd),
  tf.clip_by_norm, tf.keras.optimizers.legacy.Adam (ResourceApplyAdam's update rule), tf.range / tf.linalg.matmul /
  tf.math.ceil for the Interpolator.
Everything else resolves to an inert stub so that `TensorFlowLibrary()` (SI_Toolkit/computation_library.py:306-...) can
be constructed; none of those attributes is called on this path.  What this pins is the optimizer LOGIC of the
reference, not TensorFlow's kernels (the summation order inside reduce_mean / reduce_std is TF's own; the parity
tolerance on the distribution covers it).
"""
from __future__ import annotations

import sys
import types

import numpy as np
import torch


class _Inert:
    def __init__(self, name):
        self._n = name

    def __call__(self, *a, **k):
        return _Inert(self._n + "()")

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Inert(self._n + "." + name)

    def __mro_entries__(self, bases):
        return (object,)

    def __iter__(self):
        return iter(())


class _Mod(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        obj = _Inert(f"{self.__name__}.{name}")
        setattr(self, name, obj)
        return obj


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(dtype)
    return torch.as_tensor(np.asarray(x), dtype=dtype)


def _constant(x, dtype=None):
    a = np.asarray(x)
    if dtype is None and a.dtype.kind in "iu":  # np.tile(s, tf.constant([K, 1])) needs plain integers
        return a
    return _t(a, dtype)


def _tile(x, reps):
    return _t(x).repeat(*[int(r) for r in reps])


def _argsort(x, axis=-1, direction="ASCENDING", stable=False):
    return torch.sort(_t(x), dim=axis, descending=(direction != "ASCENDING"), stable=True).indices


def _reduce_std(x, axis=None, keepdims=False):
    x = _t(x)
    m = torch.mean(x, dim=axis, keepdim=True)
    v = torch.mean((x - m) * (x - m), dim=axis, keepdim=keepdims)
    return torch.sqrt(v)


def _variable(x, **kw):
    """tf.Variable: a leaf tensor that records gradients and can be assigned in place."""
    v = _t(x, torch.float32).detach().clone().requires_grad_(True)
    v.assign = lambda val: _assign(v, val)
    return v


def _assign(v, val):
    # a NEW storage: slices taken from the variable earlier (torch views; copies in TF) keep the values they were taken with
    v.data = _t(val).detach().clone()
    return v


class _GradientTape:
    """d(sum(target)) / d(source) by torch autograd: what tape.gradient returns for a vector target, each entry of which
    depends on one row of the source (optimizer_rpgd_tf.py:169-175)."""

    def __init__(self, **kw):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def watch(self, x):
        pass

    def gradient(self, target, source):
        return torch.autograd.grad(target.sum(), source)[0]


def _clip_by_norm(t, clip_norm, axes=None):
    """tf.clip_by_norm: t * clip_norm / max(l2norm(t, axes), clip_norm)."""
    t = _t(t)
    l2 = torch.sqrt(torch.sum(t * t, dim=tuple(axes), keepdim=True))
    c = _t(clip_norm, torch.float32)
    return t * c / torch.maximum(l2, c)


class _LegacyAdam:
    """tf.keras.optimizers.legacy.Adam for ONE variable (ResourceApplyAdam): weights = [iterations, m, v];
    lr_t = lr sqrt(1 - b2^t) / (1 - b1^t), m = b1 m + (1 - b1) g, v = b2 v + (1 - b2) g^2, var -= lr_t m / (sqrt(v) + eps)."""

    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, **kw):
        self.lr, self.b1, self.b2, self.eps = float(learning_rate), float(beta_1), float(beta_2), float(epsilon)
        self.iterations, self.m, self.v = 0, None, None

    def get_weights(self):
        if self.m is None:
            return []
        return [np.int64(self.iterations), self.m.clone(), self.v.clone()]

    def set_weights(self, w):
        if not w:
            return
        self.iterations = int(w[0])
        self.m, self.v = _t(w[1], torch.float32).clone(), _t(w[2], torch.float32).clone()

    def apply_gradients(self, grads_and_vars):
        for g, var in grads_and_vars:
            g = _t(g).detach()
            if self.m is None:
                self.m, self.v = torch.zeros_like(g), torch.zeros_like(g)
            self.iterations += 1
            t = float(self.iterations)
            lr_t = np.float32(self.lr * np.sqrt(1.0 - self.b2 ** t) / (1.0 - self.b1 ** t))
            self.m = self.b1 * self.m + (1.0 - self.b1) * g
            self.v = self.b2 * self.v + (1.0 - self.b2) * g * g
            with torch.no_grad():
                var.sub_(lr_t * self.m / (torch.sqrt(self.v) + self.eps))


def install():
    """Put the shim into sys.modules as `tensorflow` (before oracle.ref_loader.load()).  Idempotent."""
    cur = sys.modules.get("tensorflow")
    if getattr(cur, "_cps_shim", False):
        return cur
    tf = _Mod("tensorflow")
    tf.__path__ = []
    tf._cps_shim = True
    tf.float32, tf.float64, tf.int32 = torch.float32, torch.float64, torch.int32
    tf.Tensor = torch.Tensor
    tf.constant = _constant
    tf.convert_to_tensor = lambda x, dtype=None: _t(x, dtype)
    tf.zeros = lambda shape, dtype=torch.float32: torch.zeros(tuple(int(v) for v in np.atleast_1d(shape)), dtype=dtype)
    tf.ones = lambda shape, dtype=torch.float32: torch.ones(tuple(int(v) for v in np.atleast_1d(shape)), dtype=dtype)
    tf.tile = _tile
    tf.multiply = lambda a, b: _t(a) * _t(b)
    tf.clip_by_value = lambda x, lo, hi: torch.minimum(torch.maximum(_t(x), _t(lo)), _t(hi))
    tf.argsort = _argsort
    tf.gather = lambda x, idx, axis=0: torch.index_select(_t(x), axis, _t(idx).reshape(-1).long())
    tf.reduce_mean = lambda x, axis=None, keepdims=False: torch.mean(_t(x), dim=axis, keepdim=keepdims)
    tf.concat = lambda xs, axis=0: torch.cat([_t(x, torch.float32) for x in xs], dim=axis)
    tf.squeeze = lambda x, axis=None: torch.squeeze(_t(x)) if axis is None else torch.squeeze(_t(x), axis)
    tf.ensure_shape = lambda x, shape: x
    # optimizer_cem_gmm_tf (:72-93)
    tf.newaxis = None
    tf.transpose = lambda x, perm=None: _t(x).permute(*perm) if perm is not None else _t(x).t()
    tf.norm = lambda x, axis=None: torch.sqrt(torch.sum(_t(x) * _t(x), dim=tuple(axis) if isinstance(axis, (list, tuple)) else axis))
    tf.argmin = lambda x, axis=None: torch.argmin(_t(x), dim=axis)   # the first of equal minima, like TF
    tf.cast = lambda x, dtype=None: _t(x).to(dtype) if isinstance(x, torch.Tensor) else torch.as_tensor(x, dtype=dtype)
    tf.shape = lambda x: torch.as_tensor(list(_t(x).shape))
    tf.stack = lambda xs, axis=0: torch.stack([_t(x, torch.float32) for x in xs], dim=axis)
    # tf tensors hand out numpy copies whether or not a tape watches them
    if not getattr(torch.Tensor.numpy, "_cps_detaching", False):
        _orig_numpy = torch.Tensor.numpy

        def _numpy(self, *a, **k):
            return _orig_numpy(self.detach(), *a, **k).copy()   # tf hands out copies; torch would alias the variable
        _numpy._cps_detaching = True
        torch.Tensor.numpy = _numpy
    # optimizer_rpgd_tf: variables, the gradient tape, clip_by_norm, Keras' legacy Adam; Interpolator through TensorFlowLibrary
    tf.__version__ = "2.11.0"
    tf.Variable = _variable
    tf.GradientTape = _GradientTape
    tf.clip_by_norm = _clip_by_norm
    tf.zeros_like = lambda x, dtype=None: torch.zeros_like(_t(x))
    tf.range = lambda *a, **k: torch.arange(*[int(v) for v in a])
    tf.function = lambda f=None, **k: (f if f is not None else (lambda g: g))
    linalg = _Mod("tensorflow.linalg")
    linalg.matmul = lambda a, b: torch.matmul(_t(a), _t(b))
    tf.linalg = linalg
    keras = _Mod("tensorflow.keras")
    opts = _Mod("tensorflow.keras.optimizers")
    legacy = _Mod("tensorflow.keras.optimizers.legacy")
    legacy.Adam = opts.Adam = _LegacyAdam
    opts.legacy = legacy
    keras.optimizers = opts
    tf.keras = keras
    math = _Mod("tensorflow.math")
    math.reduce_std = _reduce_std
    math.ceil = lambda x: float(np.ceil(x))
    tf.math = math
    sys.modules["tensorflow"] = tf
    sys.modules["tensorflow.math"] = math
    _install_tfp()
    return tf


# ---- tensorflow_probability.python.distributions: the three classes optimizer_cem_gmm_tf builds -----------------------
GMM_DRAWS = None   # an InjectedDraws: MixtureSameFamily.sample takes one normal [n, *batch, k] and one uniform [n, *batch]


class _Normal:
    def __init__(self, loc, scale):
        self.loc, self.scale = _t(loc, torch.float32), _t(scale, torch.float32)

    def mean(self):
        return self.loc

    def stddev(self):
        return self.scale


class _Categorical:
    def __init__(self, probs):
        self.probs = torch.stack([_t(p, torch.float32).reshape(()) for p in probs]) if isinstance(probs, (list, tuple)) else _t(probs, torch.float32)


class _MixtureSameFamily:
    """Two components.  sample(): every batch member of every sample picks its component independently (index 0 iff the
    uniform draw is below probs[0]) from component draws loc + scale * eps, as tfpd.MixtureSameFamily.sample does (it
    draws all components and masks with the one-hot mixture sample)."""

    def __init__(self, mixture_distribution, components_distribution):
        self.mixture_distribution, self.components_distribution = mixture_distribution, components_distribution

    def sample(self, sample_shape):
        n = int(sample_shape[0])
        cd = self.components_distribution
        batch = list(cd.loc.shape[:-1])
        eps = GMM_DRAWS.normal([n] + batch + [int(cd.loc.shape[-1])])
        u = GMM_DRAWS.uniform([n] + batch)
        x = cd.loc[None] + cd.scale[None] * eps
        c = (u >= self.mixture_distribution.probs[0]).long()
        return torch.gather(x, -1, c[..., None])[..., 0]


def _install_tfp():
    tfp = _Mod("tensorflow_probability")
    tfp.__path__ = []
    py = _Mod("tensorflow_probability.python")
    py.__path__ = []
    d = _Mod("tensorflow_probability.python.distributions")
    d.Normal, d.Categorical, d.MixtureSameFamily, d.Distribution = _Normal, _Categorical, _MixtureSameFamily, object
    tfp.python, py.distributions = py, d
    sys.modules["tensorflow_probability"] = tfp
    sys.modules["tensorflow_probability.python"] = py
    sys.modules["tensorflow_probability.python.distributions"] = d


class InjectedDraws:
    """Replaces optimizer.rng (a tf.random.Generator): hands out pre-supplied tensors in call order."""

    def __init__(self, draws):
        self.draws = [torch.as_tensor(np.asarray(d), dtype=torch.float32) for d in draws]
        self.i = 0

    def _next(self, shape):
        d = self.draws[self.i]
        self.i += 1
        assert [int(v) for v in shape] == list(d.shape), (list(shape), d.shape)
        return d

    def normal(self, shape, mean=0.0, stddev=1.0, dtype=None, **kw):
        d = self._next(shape)
        if float(stddev) != 1.0 or float(mean) != 0.0:   # tf.random.Generator.normal(shape, mean, stddev)
            d = d * _t(stddev, torch.float32) + _t(mean, torch.float32)
        return d

    def uniform(self, shape, minval=None, maxval=None, dtype=None, **kw):
        return self._next(shape)

"""NumPy stand-ins for the slice of jax / flax.linen / optax the reference MODEL files use.
TEST INFRASTRUCTURE ONLY (never imported by the product).

Purpose: execute the reference's own source - ``attention.py``, ``track_autoencoder.py``,
``track_autoencoder_3d.py``, the loss in ``train.py`` - UNMODIFIED, where it lies under
/root/reference, to produce golden vectors (``tests/golden/make_golden_model.py``).  JAX, Flax and
optax are third-party dependencies that are absent here (the reference pins no versions:
``requirements.txt`` lists ``jax``, ``flax``, ``optax`` unpinned; API as of flax 0.8-0.10 /
jax 0.4.x).  What is restated here is therefore ONLY the published semantics of the framework
primitives the reference calls; every line of model logic (token order, masks, residual wiring,
einsum strings, quantiser, output slicing) is executed from the reference's files.

Primitives and the published behaviour they follow:
  nn.Dense               y = x @ kernel + bias, kernel [in, out]
  nn.DenseGeneral        contraction of the ``axis`` dims of x with the leading dims of kernel,
                         kernel [*in_dims, *features], bias [*features]
  nn.LayerNorm           eps 1e-6, statistics over the last axis, optional scale / bias
  nn.RMSNorm             eps 1e-6, x * rsqrt(mean(x^2)) * scale over the last axis
  nn.gelu                tanh approximation (``approximate=True`` is the flax default)
  nn.dot_product_attention   q / sqrt(d); logits [..., h, q, k]; ``where(mask, logits, finfo.min)``;
                         softmax over k; weighted sum of values
  nn.Module              dataclass-style fields, ``setup`` children named after the attribute,
                         ``@nn.compact`` children named by ``name=`` (or ``Class_i``), ``self.param``
  nn.scan                loop over ``in_axes`` of the scanned argument, outputs stacked on ``out_axes``
  nn.remat, lax.stop_gradient   identity in a forward pass
  jax.vmap               map over the leading axis
  jax.random.uniform     NOT restated (threefry): the golden generator injects the noise tensor
  optax.sigmoid_binary_cross_entropy   max(x,0) - x*z + log1p(exp(-|x|))  (= -z log s(x) - (1-z) log(1-s(x)))
  optax linear / cosine_decay / join schedules   their documented closed forms

dtypes: parameters and activations are float64 (the generator passes float64 trees), but the sinusoidal embedding is
evaluated in float32 exactly as under jax's default dtypes - float32 coordinates in, ``jnp.asarray([python floats])`` ->
float32 scales, ``int32 / n`` and ``int32 // 150.0`` -> float32, float32 sine (correctly rounded) - because its arguments reach
2**(31/3) ~ 1290 x |coordinate| and a float64 evaluation differs from the float32 one the reference performs by ~1e-4.
Every parameter leaf read through ``self.param`` is recorded so a test can assert that the
reference's own naming consumed exactly the tree the product loads.
"""
from __future__ import annotations

import dataclasses
import math
import sys
import types

import numpy as np

_stack = []          # modules whose methods are executing (innermost last)
_in_setup = []       # modules currently inside setup()
accessed = set()     # parameter paths read through Module.param
_uniform_hook = [None]


def set_uniform(fn):
    """fn(shape) -> array in [0,1): what jax.random.uniform returns while the reference runs."""
    _uniform_hook[0] = fn


# ----------------------------------------------------------------------------------------
# flax.linen.Module
# ----------------------------------------------------------------------------------------


class _JaxArray(np.ndarray):
    """Two jax behaviours NumPy does not have, attached to the arrays that flow through the reference code:

    * immutability: ``y = x; y += d`` rebinds y and leaves x alone (NumPy would write through the alias);
    * default (x64-disabled) promotion of integer arrays: ``int32 / python-number`` and ``int32 // python-float`` give
      float32, not float64.  This decides the dtype the Fourier features are evaluated in (``jnp.arange(T) / T`` and
      ``query_frame // time_scale_factor`` are concatenated with the coordinates before the embedding), and the embedding
      amplifies one float32 ulp of its argument to ~1e-4, so it has to be reproduced, not approximated.
    """

    def __iadd__(self, o):
        return np.add(self, o)

    def __isub__(self, o):
        return np.subtract(self, o)

    def __imul__(self, o):
        return np.multiply(self, o)

    def __itruediv__(self, o):
        return np.true_divide(self, o)

    def _weak(self, o):
        return self.dtype.kind in "iu" and isinstance(o, (int, float)) and not isinstance(o, bool)

    def __truediv__(self, o):
        if self._weak(o):
            return np.true_divide(np.asarray(self, np.float32), np.float32(o)).view(_JaxArray)
        return np.true_divide(self, o)

    def __floordiv__(self, o):
        if self._weak(o) and isinstance(o, float):
            return np.floor_divide(np.asarray(self, np.float32), np.float32(o)).view(_JaxArray)
        return np.floor_divide(self, o)


def _imm(x):
    return x.view(_JaxArray) if isinstance(x, np.ndarray) else x


def _jnp_arange(*a, **k):
    out = np.arange(*a, **k)
    return (out.astype(np.int32) if out.dtype.kind == "i" else out.astype(np.float32)).view(_JaxArray)


def _jnp_array(obj, dtype=None, **k):
    """jnp.array / jnp.asarray: Python floats become float32, Python ints int32 (jax defaults); arrays keep their dtype."""
    if dtype is None and not isinstance(obj, np.ndarray):
        probe = np.asarray(obj)
        dtype = np.float32 if probe.dtype == np.float64 else (np.int32 if probe.dtype == np.int64 else probe.dtype)
    return np.array(obj, dtype=dtype, **k).view(_JaxArray)


def _jnp_sin(x):
    """float32 sine = the correctly rounded one (XLA's is within 1 ulp of it; oracle/model.py uses the same representative)."""
    x = np.asarray(x)
    if x.dtype == np.float32:
        return np.sin(x.astype(np.float64)).astype(np.float32)
    return np.sin(x)


def _wrap(fn, is_setup=False):
    def method(self, *a, **k):
        a = tuple(_imm(x) for x in a)
        k = {n: _imm(v) for n, v in k.items()}
        if not is_setup:
            self._ensure_setup()
        _stack.append(self)
        try:
            return fn(self, *a, **k)
        finally:
            _stack.pop()

    method.__name__ = getattr(fn, "__name__", "method")
    method.__wrapped__ = fn
    return method


class Module:
    name = None

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        names = []
        for klass in reversed(cls.__mro__):
            for k in klass.__dict__.get("__annotations__", {}):
                if k not in ("name", "parent") and k not in names:
                    names.append(k)
        cls._field_names = tuple(names)
        for k, v in list(cls.__dict__.items()):
            if isinstance(v, types.FunctionType) and (not k.startswith("_") or k == "__call__"):
                setattr(cls, k, _wrap(v, is_setup=(k == "setup")))

    def __init__(self, *args, name=None, parent=None, **kw):
        d = self.__dict__
        d["_children_count"] = {}
        d["_setup_done"] = False
        d["_params"] = None
        for k, v in zip(self._field_names, args):
            kw[k] = v
        for k in self._field_names:
            if k in kw:
                d[k] = kw.pop(k)
            elif not hasattr(type(self), k):
                raise TypeError(f"{type(self).__name__}: missing field {k}")
        if kw:
            raise TypeError(f"{type(self).__name__}: unexpected fields {sorted(kw)}")
        d["name"] = name
        d["_parent"] = parent if parent is not None else (_stack[-1] if _stack else None)

    # -- naming / parameter scope --
    def __setattr__(self, key, value):
        if isinstance(value, Module) and _in_setup and _in_setup[-1] is self and value.name is None:
            value.__dict__["name"] = key
            value.__dict__["_parent"] = self
        self.__dict__[key] = value

    def __getattr__(self, key):
        if key.startswith("_") or self.__dict__.get("_setup_done", True):
            raise AttributeError(key)
        self._ensure_setup()
        try:
            return self.__dict__[key]
        except KeyError:
            raise AttributeError(key) from None

    def _ensure_setup(self):
        if not self._setup_done:
            self.__dict__["_setup_done"] = True
            if hasattr(self, "setup"):
                _in_setup.append(self)
                try:
                    self.setup()
                finally:
                    _in_setup.pop()

    def _resolve_name(self):
        if self.name is None:
            p = self._parent
            base = type(self).__name__
            i = p._children_count.get(base, 0) if p is not None else 0
            if p is not None:
                p._children_count[base] = i + 1
            self.__dict__["name"] = f"{base}_{i}"
        return self.name

    def _path(self):
        if self._parent is None:
            return ()
        return self._parent._path() + (self._resolve_name(),)

    def _scope(self):
        if self._parent is None:
            if self._params is None:
                raise RuntimeError("top-level module is not bound; use flax_shim.bind(module, params)")
            return self._params
        tree = self._parent._scope()
        n = self._resolve_name()
        if n not in tree:
            raise KeyError(f"no parameters for {'/'.join(self._path())}")
        return tree[n]

    def param(self, name, init_fn, *init_args):
        tree = self._scope()
        if name not in tree:
            raise KeyError(f"missing parameter {'/'.join(self._path() + (name,))}")
        value = np.asarray(tree[name])
        if init_args and tuple(init_args[0]) != value.shape:
            raise ValueError(f"{'/'.join(self._path() + (name,))}: tree has {value.shape}, module wants {tuple(init_args[0])}")
        accessed.add("/".join(self._path() + (name,)))
        return value


def bind(module, params):
    """The equivalent of ``module.bind({'params': params})``."""
    module.__dict__["_params"] = params
    module.__dict__["_parent"] = None
    return module


def compact(fn):
    return fn


def remat(fn, **kw):
    return fn


class Dense(Module):
    features: int
    use_bias: bool = True

    def __call__(self, x):
        kernel = self.param("kernel", None, (x.shape[-1], self.features))
        y = x @ kernel
        if self.use_bias:
            y = y + self.param("bias", None, (self.features,))
        return y


class DenseGeneral(Module):
    features: object
    axis: object = -1
    use_bias: bool = True

    def __call__(self, x):
        feats = tuple(self.features) if isinstance(self.features, (tuple, list)) else (self.features,)
        axes = tuple(self.axis) if isinstance(self.axis, (tuple, list)) else (self.axis,)
        axes = tuple(a % x.ndim for a in axes)
        in_dims = tuple(x.shape[a] for a in axes)
        kernel = self.param("kernel", None, in_dims + feats)
        y = np.tensordot(x, kernel, axes=(axes, tuple(range(len(axes)))))
        if self.use_bias:
            y = y + self.param("bias", None, feats)
        return y


class LayerNorm(Module):
    epsilon: float = 1e-6
    use_bias: bool = True
    use_scale: bool = True

    def __call__(self, x):
        mean = x.mean(-1, keepdims=True)
        var = np.maximum((x * x).mean(-1, keepdims=True) - mean * mean, 0.0)
        y = (x - mean) / np.sqrt(var + self.epsilon)
        if self.use_scale:
            y = y * self.param("scale", None, (x.shape[-1],))
        if self.use_bias:
            y = y + self.param("bias", None, (x.shape[-1],))
        return y


class RMSNorm(Module):
    epsilon: float = 1e-6
    use_scale: bool = True

    def __call__(self, x):
        y = x / np.sqrt((x * x).mean(-1, keepdims=True) + self.epsilon)
        if self.use_scale:
            y = y * self.param("scale", None, (x.shape[-1],))
        return y


def gelu(x, approximate=True):
    if approximate:
        return 0.5 * x * (1.0 + np.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))
    erf = np.vectorize(math.erf)
    return 0.5 * x * (1.0 + erf(x / math.sqrt(2.0)))


def softmax(x, axis=-1):
    e = np.exp(x - x.max(axis=axis, keepdims=True))
    return e / e.sum(axis=axis, keepdims=True)


def dot_product_attention(query, key, value, bias=None, mask=None, **unused):
    depth = query.shape[-1]
    logits = np.einsum("...qhd,...khd->...hqk", query / np.sqrt(depth).astype(query.dtype), key)
    if bias is not None:
        logits = logits + bias
    if mask is not None:
        logits = np.where(mask, logits, np.finfo(logits.dtype).min)
    weights = softmax(logits, -1)
    return np.einsum("...hqk,...khd->...qhd", weights, value)


def scan(fn, variable_broadcast=None, split_rngs=None, in_axes=0, out_axes=0, **unused):
    def run(module, carry, xs):
        outs = []
        for i in range(xs.shape[in_axes]):
            carry, y = fn(module, carry, np.take(xs, i, axis=in_axes))
            outs.append(y)
        return carry, tree_map(lambda *leaves: np.stack(leaves, axis=out_axes), *outs)

    return run


# ----------------------------------------------------------------------------------------
# jax / jax.numpy / flax.struct / optax
# ----------------------------------------------------------------------------------------


def tree_map(f, tree, *rest):
    if dataclasses.is_dataclass(tree) and not isinstance(tree, type):
        return type(tree)(**{fl.name: tree_map(f, getattr(tree, fl.name), *[getattr(r, fl.name) for r in rest])
                             for fl in dataclasses.fields(tree)})
    if isinstance(tree, dict):
        return {k: tree_map(f, v, *[r[k] for r in rest]) for k, v in tree.items()}
    if isinstance(tree, (list, tuple)):
        return type(tree)(tree_map(f, v, *[r[i] for r in rest]) for i, v in enumerate(tree))
    return f(tree, *rest)


def vmap(f, **unused):
    return lambda x: np.stack([f(xi) for xi in x])


def struct_dataclass(cls=None, **kw):
    wrap = lambda c: dataclasses.dataclass(frozen=True)(c)
    return wrap(cls) if cls is not None else wrap


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def sigmoid_binary_cross_entropy(logits, labels):
    return np.maximum(logits, 0.0) - logits * labels + np.log1p(np.exp(-np.abs(logits)))


def linear_schedule(init_value, end_value, transition_steps, transition_begin=0):
    def f(count):
        c = np.clip(count - transition_begin, 0, transition_steps)
        return init_value + (end_value - init_value) * c / transition_steps
    return f


def cosine_decay_schedule(init_value, decay_steps, alpha=0.0, exponent=1.0):
    def f(count):
        c = np.minimum(count, decay_steps)
        cosine = 0.5 * (1.0 + np.cos(np.pi * c / decay_steps))
        return init_value * ((1.0 - alpha) * cosine ** exponent + alpha)
    return f


def join_schedules(schedules, boundaries):
    def f(step):
        out = schedules[0](step)
        for b, s in zip(boundaries, schedules[1:]):
            out = np.where(step < b, out, s(step - b))
        return out
    return f


class _Namespace(types.ModuleType):
    """A module whose missing attributes fall through to another module (numpy)."""

    def __init__(self, name, fallback=None, **attrs):
        super().__init__(name)
        self.__dict__.update(attrs)
        self.__dict__["_fallback"] = fallback

    def __getattr__(self, item):
        fb = self.__dict__.get("_fallback")
        if fb is not None and hasattr(fb, item):
            return getattr(fb, item)
        raise AttributeError(f"flax_shim: {self.__name__}.{item} is not provided")


def _uniform(key, shape, *a, **k):
    if _uniform_hook[0] is None:
        raise RuntimeError("jax.random.uniform is not restated; inject the noise with flax_shim.set_uniform")
    return np.asarray(_uniform_hook[0](tuple(shape)))


def modules():
    """Fresh stand-in modules keyed by import name."""
    jnp = _Namespace("jax.numpy", fallback=np, arange=_jnp_arange, array=_jnp_array, asarray=_jnp_array, sin=_jnp_sin)
    jnn = _Namespace("jax.nn", sigmoid=sigmoid, softmax=softmax, gelu=gelu)
    random = _Namespace("jax.random", PRNGKey=lambda seed: ("key", seed), uniform=_uniform)
    lax = _Namespace("jax.lax", stop_gradient=lambda x: x)
    tree_util = _Namespace("jax.tree_util", tree_map=tree_map)
    jax = _Namespace("jax", numpy=jnp, nn=jnn, random=random, lax=lax, tree_util=tree_util, vmap=vmap,
                     jit=lambda f=None, **k: (f if f is not None else (lambda g: g)))
    initializers = _Namespace("flax.linen.initializers", normal=lambda stddev=1e-2, **k: (lambda *a, **kk: None))
    linen = _Namespace("flax.linen", Module=Module, compact=compact, remat=remat, scan=scan, Dense=Dense,
                       DenseGeneral=DenseGeneral, LayerNorm=LayerNorm, RMSNorm=RMSNorm, gelu=gelu,
                       dot_product_attention=dot_product_attention, initializers=initializers)
    struct = _Namespace("flax.struct", dataclass=struct_dataclass)
    flax = _Namespace("flax", linen=linen, struct=struct)
    optax = _Namespace("optax", sigmoid_binary_cross_entropy=sigmoid_binary_cross_entropy, linear_schedule=linear_schedule,
                       cosine_decay_schedule=cosine_decay_schedule, join_schedules=join_schedules)
    return {"jax": jax, "jax.numpy": jnp, "jax.nn": jnn, "jax.random": random, "jax.lax": lax, "jax.tree_util": tree_util,
            "flax": flax, "flax.linen": linen, "flax.linen.initializers": initializers, "flax.struct": struct, "optax": optax}


_loaded = {}


def load_reference(reference_dir="/root/reference"):
    """Execute attention.py, track_autoencoder.py, track_autoencoder_3d.py where they lie, on the stand-ins.
    Returns {'attention': module, 'track_autoencoder': module, 'track_autoencoder_3d': module}."""
    if _loaded:
        return _loaded
    import importlib.util
    import os

    names = ("attention", "track_autoencoder", "track_autoencoder_3d")
    stubs = modules()
    saved = {k: sys.modules.get(k) for k in list(stubs) + list(names)}
    sys.modules.update(stubs)
    try:
        for n in names:
            spec = importlib.util.spec_from_file_location(n, os.path.join(reference_dir, n + ".py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[n] = mod   # track_autoencoder imports `attention` by bare name
            spec.loader.exec_module(mod)
            _loaded[n] = mod
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return _loaded


def load_loss(reference_dir="/root/reference"):
    """compute_loss_3d / compute_loss_2d / create_learning_rate_schedule of train.py (train.py:41-129): the functions are
    re-executed from the reference file's own source text, cut at the first line that needs the absent training stack."""
    import os

    src = open(os.path.join(reference_dir, "train.py")).read()
    start = src.index("def create_learning_rate_schedule")
    stop = src.index("@functools.partial(jax.jit")
    stubs = modules()
    env = {"jnp": stubs["jax.numpy"], "jax": stubs["jax"], "optax": stubs["optax"], "np": np}
    exec(compile(src[start:stop], os.path.join(reference_dir, "train.py"), "exec"), env)
    return env

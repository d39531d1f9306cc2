"""Import the reference's pure-NumPy functions UNMODIFIED.  TEST INFRASTRUCTURE.

Only usable where /root/reference exists (the build container, not the GPU box).  The
reference's ``inference.py`` / ``data_loader.py`` import jax/flax/tensorflow at module scope,
none of which are installed; the functions we need (``lift_2d_to_3d``,
``sample_dino_features_for_tracks``, ``sample_depth_features_for_tracks``,
``_unflatten_params``, ``load_checkpoint`` .npz branches, ``prepare_3d_batch``) are pure
NumPy.  We satisfy the imports with inert stub modules in ``sys.modules`` and execute the
reference files where they lie.  Used by ``tests/golden/make_golden.py`` (fixtures) and by
``bench.py --impl reference`` when the directory is present.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE_DIR = os.environ.get("SPA3D_REFERENCE_DIR", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "inference.py"))


class _Anything:
    """Attribute sink: any attribute access / call returns another sink or a passthrough."""

    def __init__(self, name="stub"):
        self._name = name

    def __getattr__(self, item):
        return _Anything(f"{self._name}.{item}")

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]  # decorator use (@nn.compact, @nn.remat, @jax.jit ...)
        return _Anything(self._name + "()")

    def __mro_entries__(self, bases):
        return (object,)


def _stub_module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)

    def _getattr(item, _n=name):
        if item.startswith("__"):
            raise AttributeError(item)
        return _Anything(f"{_n}.{item}")

    m.__getattr__ = _getattr  # type: ignore[attr-defined]
    return m


def _install_stubs():
    class _Module:  # flax.linen.Module stand-in (classes are defined, never instantiated)
        pass

    def _dataclass(cls=None, **k):
        return cls if cls is not None else (lambda c: c)

    jnp = _stub_module("jax.numpy", array=np.array, asarray=np.asarray, float32=np.float32, int32=np.int32)
    jax = _stub_module("jax", numpy=jnp)
    jax.nn = _stub_module("jax.nn")
    linen = _stub_module("flax.linen", Module=_Module, compact=lambda f: f, remat=lambda f: f)
    struct = _stub_module("flax.struct", dataclass=_dataclass)
    training = _stub_module("flax.training")
    ckpt = _stub_module("flax.training.checkpoints")
    training.checkpoints = ckpt
    flax = _stub_module("flax", linen=linen, struct=struct, training=training)
    stubs = {
        "jax": jax,
        "jax.numpy": jnp,
        "jax.nn": jax.nn,
        "flax": flax,
        "flax.linen": linen,
        "flax.struct": struct,
        "flax.training": training,
        "flax.training.checkpoints": ckpt,
        "tensorflow": _stub_module("tensorflow"),
        "tensorflow_datasets": _stub_module("tensorflow_datasets"),
        "torchvision": _stub_module("torchvision"),
        "torchvision.transforms": _stub_module("torchvision.transforms"),
        "optax": _stub_module("optax"),
        "chex": _stub_module("chex"),
        "wandb": _stub_module("wandb"),
        "transformers": _stub_module("transformers"),
        "tapnet": _stub_module("tapnet"),
        "tapnet.tapvid3d": _stub_module("tapnet.tapvid3d"),
        "tapnet.tapvid3d.evaluation": _stub_module("tapnet.tapvid3d.evaluation"),
        "tapnet.tapvid3d.splits": _stub_module("tapnet.tapvid3d.splits"),
    }
    # absl: always inert - the reference scripts re-define the same flag names, which the real absl rejects
    for name in ("absl", "absl.app", "absl.flags", "absl.logging"):
        stubs[name] = _stub_module(name)
    for opt in ("cv2", "einops"):
        try:
            __import__(opt)
        except Exception:  # pragma: no cover
            stubs[opt] = _stub_module(opt)
    saved = {}
    for k, v in stubs.items():
        saved[k] = sys.modules.get(k)
        sys.modules[k] = v
    return saved


def _restore(saved):
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


_cache = {}


def load(module_name: str):
    """Execute /root/reference/<module_name>.py with stubbed heavy imports; return module."""
    if module_name in _cache:
        return _cache[module_name]
    if not available():
        raise RuntimeError(f"reference sources not present at {REFERENCE_DIR}")
    import torch  # noqa: F401  (real torch must be imported before the stubs go in)

    saved = _install_stubs()
    sys.path.insert(0, REFERENCE_DIR)
    try:
        spec = importlib.util.spec_from_file_location(
            f"_ref_{module_name}", os.path.join(REFERENCE_DIR, module_name + ".py")
        )
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(REFERENCE_DIR)
        for name in ("attention", "track_autoencoder", "track_autoencoder_3d"):
            sys.modules.pop(name, None)
        _restore(saved)
    _cache[module_name] = mod
    return mod

"""TEST INFRASTRUCTURE ONLY -- runs the reference's pure-JAX ``HashGridEncoder`` (models/encoders.py:58-256,
UNMODIFIED, imported from /root/reference) without jax/flax, to pin the CPU oracle to the reference's own code.

jax, jax.numpy, flax.linen, chex, jaxtcnn and shjax are absent from this image.  This module installs just enough of
them for ``models/encoders.py`` to import and for ``HashGridEncoder.__call__`` to run eagerly on numpy arrays:

* ``jax.numpy`` -> numpy, with jax's default dtypes (Python floats -> float32, ints -> int32; x64 is off in the
  reference) and uint32 reductions that stay uint32;
* ``jax.vmap`` -> a Python loop over the mapped axis (the encoder vmaps tiny per-point functions over [L, n]);
* ``flax.linen.Module`` -> a dataclass base whose ``self.param(name, init, shape, dtype)`` hands back an array bound
  beforehand (``bind_params``), shape- and dtype-checked; ``nn.compact`` -> identity;
* ``utils.common.{next_multiple, vmap_jaxfn_with, jit_jaxfn_with}`` and ``utils.types.empty_impl``: the reference's
  OWN source of those functions, extracted by AST and executed here (their modules import tensorflow, git, ...).

Only oracle/make_golden_encoder.py uses this, in the container that has /root/reference; the vectors it writes are
committed under tests/golden/.  Never imported by the product, the tests or the bench.
"""
import ast
import dataclasses
import functools
import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE = "/root/reference"


def _asarray(x, dtype=None):
    a = np.asarray(x, dtype=dtype)
    if dtype is None:  # jax defaults with x64 disabled
        if a.dtype == np.float64:
            a = a.astype(np.float32)
        elif a.dtype == np.int64:
            a = a.astype(np.int32)
    return a


def _sum(a, axis=None, **kw):
    a = np.asarray(a)
    return np.sum(a, axis=axis, dtype=a.dtype, **kw)  # jax keeps uint32 / float32 accumulators


def vmap(fun=None, in_axes=0, out_axes=0, axis_name=None, axis_size=None, spmd_axis_name=None):
    if fun is None:
        return functools.partial(vmap, in_axes=in_axes, out_axes=out_axes)

    @functools.wraps(fun)
    def mapped(*args):
        axes = tuple(in_axes) if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = next(np.shape(a)[ax] for a, ax in zip(args, axes) if ax is not None)
        outs = [fun(*[a if ax is None else np.take(a, i, axis=ax) for a, ax in zip(args, axes)]) for i in range(n)]
        if isinstance(outs[0], tuple):
            return tuple(np.stack([o[k] for o in outs], axis=out_axes) for k in range(len(outs[0])))
        return np.stack(outs, axis=out_axes)

    return mapped


class _Stub(types.ModuleType):
    """A module whose unknown attributes are inert placeholders (annotations, unused imports)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        value = type(name, (), {"__init__": lambda self, *a, **k: None, "__call__": lambda self, *a, **k: None})
        setattr(self, name, value)
        return value


def _numpy_namespace(name):
    m = _Stub(name)
    for k in dir(np):
        if not k.startswith("_"):
            setattr(m, k, getattr(np, k))
    m.asarray = _asarray
    m.array = _asarray
    m.sum = _sum
    return m


class Module:
    """flax.linen.Module stand-in: subclasses become dataclasses (as flax makes them), parameters are pre-bound."""

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        dataclasses.dataclass(cls)

    def bind_params(self, **named):
        object.__setattr__(self, "_bound_params", dict(named))
        return self

    def param(self, name, init_fn, shape, dtype=np.float32):
        value = getattr(self, "_bound_params", {}).get(name)
        if value is None:
            raise KeyError(f"parameter {name!r} not bound (reference asks for shape {shape}, dtype {dtype})")
        if tuple(value.shape) != tuple(shape) or value.dtype != np.dtype(dtype):
            raise ValueError(f"parameter {name!r}: reference expects {tuple(shape)} {np.dtype(dtype)}, got "
                             f"{tuple(value.shape)} {value.dtype}")
        return value


def _extract_functions(path, names):
    """The reference's own source of a few small pure-Python helpers, executed against the shim."""
    tree = ast.parse(open(path).read())
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    missing = set(names) - {n.name for n in keep}
    if missing:
        raise RuntimeError(f"{path}: {sorted(missing)} not found")
    return ast.Module(body=keep, type_ignores=[])


def install():
    """Put the stand-ins into sys.modules and return the reference's ``models.encoders`` module."""
    jnp = _numpy_namespace("jax.numpy")
    jax = _Stub("jax")
    jax.numpy, jax.vmap, jax.Array = jnp, vmap, np.ndarray
    jax.jit = lambda fun=None, **kw: fun if fun is not None else (lambda f: f)
    jax.random = _Stub("jax.random")
    src_lib = _Stub("jax._src.lib")
    linen = _Stub("flax.linen")
    linen.Module, linen.compact = Module, (lambda f: f)
    dtypes = _Stub("flax.linen.dtypes")
    dtypes.Dtype = object
    flax = _Stub("flax")
    flax.linen = linen
    mods = {"jax": jax, "jax.numpy": jnp, "jax.random": jax.random, "jax._src": _Stub("jax._src"), "jax._src.lib": src_lib,
            "flax": flax, "flax.linen": linen, "flax.linen.dtypes": dtypes, "chex": _Stub("chex"),
            "jaxtcnn": _Stub("jaxtcnn"), "shjax": _Stub("shjax")}
    # utils.common / utils.types: only the helpers encoders.py imports, from the reference's own source
    utils = types.ModuleType("utils")
    utils.__path__ = []
    common = types.ModuleType("utils.common")
    import typing
    common.__dict__.update(functools=functools, jax=jax, Any=typing.Any, Hashable=typing.Hashable, Sequence=typing.Sequence,
                           Iterable=typing.Iterable, xc=src_lib)
    exec(compile(_extract_functions(os.path.join(REFERENCE, "utils", "common.py"),
                                    {"next_multiple", "vmap_jaxfn_with", "jit_jaxfn_with"}), "utils/common.py", "exec"),
         common.__dict__)
    rtypes = types.ModuleType("utils.types")
    rtypes.__dict__.update(Type=typing.Type)
    exec(compile(_extract_functions(os.path.join(REFERENCE, "utils", "types.py"), {"empty_impl"}), "utils/types.py", "exec"),
         rtypes.__dict__)
    mods.update({"utils": utils, "utils.common": common, "utils.types": rtypes})
    sys.modules.update(mods)
    spec = importlib.util.spec_from_file_location("reference_models_encoders", os.path.join(REFERENCE, "models", "encoders.py"))
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module

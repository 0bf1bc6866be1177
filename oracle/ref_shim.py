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


# ----------------------------------------------------------------------------------------------------------------------
# Density-grid update (utils/types.py:93-144 OccupancyDensityGrid, :1149-1239 NeRFState.update_ogrid_density /
# threshold_ogrid): the reference's own source of that class and of those methods, extracted by AST and executed on
# numpy.  utils/types.py cannot be imported (flax.training, pydantic, PIL, the compiled volrendjax ops ...), and
# NeRFState is a flax TrainState; the methods only touch a handful of attributes, which `RefStateBase` provides.
def _down(x):
    """jax with x64 disabled never yields 64-bit values: int (+) Python float is float32, mixed int/float32 is float32."""
    if isinstance(x, np.ndarray):
        if x.dtype == np.float64:
            return x.astype(np.float32)
        if x.dtype == np.int64:
            return x.astype(np.int32)
        if x.dtype == np.uint64:
            return x.astype(np.uint32)
    return x


class JArr(np.ndarray):
    """numpy array with jax's functional update syntax (``a.at[idx].set(v)`` returns an updated copy) and jax's
    32-bit result types for arithmetic."""

    @property
    def at(self):
        return _At(self)

    def __getitem__(self, idx):
        # jax clamps out-of-bounds gather indices (idle renderer slots carry index >= n_total_rays)
        first = idx[0] if isinstance(idx, tuple) else idx
        if isinstance(first, np.ndarray) and first.dtype.kind in "iu" and first.ndim == 1 and self.shape:
            clamped = np.minimum(np.asarray(first).astype(np.int64), self.shape[0] - 1)
            idx = (clamped,) + tuple(idx[1:]) if isinstance(idx, tuple) else clamped
        return super().__getitem__(idx)

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        args = [np.asarray(i) if isinstance(i, JArr) else i for i in inputs]
        if out is not None:
            kwargs["out"] = tuple(np.asarray(o) if isinstance(o, JArr) else o for o in out)
        r = getattr(ufunc, method)(*args, **kwargs)
        if out is not None:
            return out[0] if len(out) == 1 else out
        return tuple(_j(_down(x)) for x in r) if isinstance(r, tuple) else _j(_down(r))


class _At:
    def __init__(self, a):
        self.a = a

    def __getitem__(self, idx):
        return _AtIdx(self.a, idx)


class _AtIdx:
    def __init__(self, a, idx):
        self.a, self.idx = a, idx

    def set(self, v):
        out = self.a.copy()
        idx = self.idx
        if isinstance(idx, np.ndarray) and idx.dtype.kind in "iu" and idx.ndim == 1:
            ok = np.asarray(idx).astype(np.int64) < out.shape[0]  # jax drops out-of-bounds scatter indices
            if not ok.all():
                np.asarray(out)[np.asarray(idx)[ok]] = np.asarray(v)[ok] if np.ndim(v) else v
                return out
        out[idx] = v
        return out


def _j(x):
    return x.view(JArr) if isinstance(x, np.ndarray) and not isinstance(x, JArr) else x


class _JnpForTypes(types.ModuleType):
    """jax.numpy for the extracted code: numpy functions whose array results carry ``.at``."""

    def __getattr__(self, name):
        target = {"asarray": _asarray, "array": _asarray, "sum": _sum,
                  "clip": lambda a, a_min=None, a_max=None: np.clip(a, a_min, a_max)}.get(name) or getattr(np, name)
        if not callable(target) or isinstance(target, type):
            return target

        @functools.wraps(target)
        def wrapped(*a, **k):
            r = target(*a, **k)
            if isinstance(r, (list, tuple)):
                return type(r)(_j(_down(x)) for x in r)
            return _j(_down(r))

        return wrapped


class ScriptedRandom(types.ModuleType):
    """jax.random with the draws supplied by the caller (the oracle takes them as inputs too): ``choice`` hands out
    the scripted index arrays in call order, ``uniform`` maps scripted U[0,1) numbers onto [minval, maxval) with
    jax.random.uniform's arithmetic (u * (maxval - minval) + minval, clamped below at minval)."""

    KeyArray = object

    def __init__(self, choices, uniforms):
        super().__init__("jax.random")
        self._choices, self._uniforms = list(choices), list(uniforms)

    def split(self, key, num=2):
        return [None] * num

    def choice(self, key, a, shape, replace=True, p=None):
        draw = self._choices.pop(0)
        assert tuple(draw.shape) == tuple(shape), (draw.shape, shape)
        if p is not None:  # the scripted cells must be ones the reference could have drawn
            pos = {int(v): i for i, v in enumerate(np.asarray(a))}
            assert all(np.asarray(p)[pos[int(v)]] > 0 for v in draw[:64])
        return _j(np.asarray(draw, dtype=np.asarray(a).dtype))

    def uniform(self, key, shape, dtype, minval=0.0, maxval=1.0):
        u = np.asarray(self._uniforms.pop(0), np.float32)
        assert tuple(u.shape) == tuple(shape)
        lo, hi = np.float32(minval), np.float32(maxval)
        return np.maximum(lo, u * (hi - lo) + lo).astype(dtype)


class RefStateBase:
    """The attributes NeRFState's grid-update methods read (utils/types.py:1149-1239), and ``replace``."""

    def __init__(self, ogrid, nerf_fn, G, diagonal_n_steps, bound, cascades):
        self.ogrid, self.nerf_fn = ogrid, nerf_fn
        self.raymarch = types.SimpleNamespace(density_grid_res=G, diagonal_n_steps=diagonal_n_steps)
        self.scene_meta = types.SimpleNamespace(bound=bound, cascades=cascades)
        self.locked_params = {"nerf": None}

    def replace(self, **kw):
        new = object.__new__(type(self))
        new.__dict__.update(self.__dict__)
        new.__dict__.update(kw)
        return new


class _Progress:
    """tqdm stand-in: iterates, swallows the progress-bar calls."""

    def __init__(self, iterable, **kw):
        self._it = iterable

    def __iter__(self):
        return iter(self._it)

    def set_description_str(self, *a, **k):
        pass

    def close(self):
        pass


def install_grid_update(oracle_module, jran, extra_methods=()):
    """Returns (OccupancyDensityGrid, RefState): the reference's class and a state class carrying the reference's
    ``update_ogrid_density`` / ``threshold_ogrid`` / ``density_threshold_from_min_step_size`` (plus ``extra_methods``
    of NeRFState, e.g. ``mark_untrained_density_grid``).  The two compiled ops
    they call (volrendjax.morton3d_invert, volrendjax.packbits) are served by the C oracle, itself pinned to the
    reference's CUDA kernels by tests/golden/morton_packbits.npz."""
    import typing
    path = os.path.join(REFERENCE, "utils", "types.py")
    tree = ast.parse(open(path).read())
    jnp = _JnpForTypes("jax.numpy")
    jax = _Stub("jax")
    jax.Array, jax.numpy, jax.random = np.ndarray, jnp, jran
    jax.jit = lambda fun=None, **kw: fun if fun is not None else (lambda f: f)
    jax.lax = types.SimpleNamespace(stop_gradient=lambda x: x)

    def flax_dataclass(cls):  # flax.struct.dataclass: a frozen dataclass with .replace
        cls = dataclasses.dataclass(cls)
        cls.replace = lambda self, **kw: dataclasses.replace(self, **kw)
        return cls

    struct = types.SimpleNamespace(field=lambda pytree_node=True, **kw: dataclasses.field(**kw))

    def packbits(density_threshold, density_grid):
        mask, bits = oracle_module.packbits(float(density_threshold), np.asarray(density_grid, np.float32))
        return _j(np.asarray(mask).astype(bool)), _j(np.asarray(bits))

    ns = dict(jax=jax, jnp=jnp, jran=jran, np=np, List=typing.List, struct=struct, dataclass=flax_dataclass,
              morton3d_invert=lambda idx: np.asarray(oracle_module.morton3d_invert(np.asarray(idx, np.uint32))),
              packbits=packbits, RefStateBase=RefStateBase, tqdm=_Progress, tqdm_format="")
    exec(compile(_extract_functions(path, {"empty_impl"}), "utils/types.py", "exec"), ns)
    grid_cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "OccupancyDensityGrid")
    exec(compile(ast.Module(body=[grid_cls], type_ignores=[]), "utils/types.py", "exec"), ns)
    state_cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "NeRFState")
    wanted = {"update_ogrid_density", "threshold_ogrid", "density_threshold_from_min_step_size"} | set(extra_methods)
    methods = [n for n in state_cls.body if isinstance(n, ast.FunctionDef) and n.name in wanted]
    assert {m.name for m in methods} == wanted
    ref_state = ast.ClassDef(name="RefState", bases=[ast.Name(id="RefStateBase", ctx=ast.Load())], keywords=[], body=methods,
                             decorator_list=[], type_params=[])
    module = ast.Module(body=[ref_state], type_ignores=[])
    ast.fix_missing_locations(module)
    exec(compile(module, "utils/types.py", "exec"), ns)
    return ns["OccupancyDensityGrid"], ns["RefState"]


def install_rays():
    """The reference's ray generation, unmodified: ``make_rays_worldspace`` / ``make_near_far_from_bound``
    (models/renderers/cuda.py:22-97) and ``Camera.make_ray_directions_from_pixel_coordinates`` (utils/types.py:398-439)
    for an undistorted perspective camera, plus the per-pixel ray construction nested in ``train_step``
    (app/nerf/_utils.py:96-115)."""
    import typing
    jnp = _JnpForTypes("jax.numpy")
    jnp.pi = np.float32(np.pi)
    jnp.linalg = types.SimpleNamespace(norm=lambda a, **k: _j(_down(np.linalg.norm(np.asarray(a), **k))))
    jax = _Stub("jax")
    jax.Array, jax.numpy = np.ndarray, jnp
    jax.jit = lambda fun=None, **kw: fun if fun is not None else (lambda f: f)
    chex = _Stub("chex")
    for name in ("assert_type", "assert_rank", "assert_equal_shape"):
        setattr(chex, name, lambda *a, **k: None)
    ns = dict(jax=jax, jnp=jnp, chex=chex, np=np, Tuple=typing.Tuple, Camera=object, RigidTransformation=object)
    types_tree = ast.parse(open(os.path.join(REFERENCE, "utils", "types.py")).read())
    cam_cls = next(n for n in types_tree.body if isinstance(n, ast.ClassDef) and n.name == "Camera")
    method = next(n for n in cam_cls.body if isinstance(n, ast.FunctionDef) and n.name == "make_ray_directions_from_pixel_coordinates")
    ref_cam = ast.ClassDef(name="RefCamera", bases=[], keywords=[], body=[method], decorator_list=[], type_params=[])
    module = ast.Module(body=[ref_cam], type_ignores=[])
    ast.fix_missing_locations(module)
    exec(compile(module, "utils/types.py", "exec"), ns)
    exec(compile(_extract_functions(os.path.join(REFERENCE, "models", "renderers", "cuda.py"),
                                    {"make_rays_worldspace", "make_near_far_from_bound"}), "models/renderers/cuda.py", "exec"), ns)

    # the ray construction inside train_step (app/nerf/_utils.py:96-115) is a nested function over the enclosing
    # scope's `scene`, `view_idcs`, `pixel_idcs`: compiled at module level those become globals of `train_ns`
    utils_tree = ast.parse(open(os.path.join(REFERENCE, "app", "nerf", "_utils.py")).read())
    train_step = next(n for n in utils_tree.body if isinstance(n, ast.FunctionDef) and n.name == "train_step")
    nested = next(n for n in train_step.body if isinstance(n, ast.FunctionDef) and n.name == "make_rays_worldspace")
    train_ns = dict(jnp=jnp, jax=jax, Tuple=typing.Tuple)
    exec(compile(ast.Module(body=[nested], type_ignores=[]), "app/nerf/_utils.py", "exec"), train_ns)

    def train_rays(camera, transforms, perm):
        """view/pixel split as SceneData.get_view_indices / get_pixel_indices (utils/types.py:1029-1039), then the
        reference's nested make_rays_worldspace."""
        perm = _j(np.asarray(perm, np.uint32))
        train_ns["view_idcs"] = jnp.floor_divide(perm, camera.n_pixels)
        train_ns["pixel_idcs"] = jnp.mod(perm, camera.n_pixels)
        train_ns["scene"] = types.SimpleNamespace(meta=types.SimpleNamespace(camera=camera), transforms=_j(np.asarray(transforms, np.float32)))
        return train_ns["make_rays_worldspace"]()

    def make_camera(width, height, fx, fy, cx, cy):
        cam = ns["RefCamera"]()
        cam.__dict__.update(width=width, height=height, n_pixels=width * height, fx=fx, fy=fy, cx=cx, cy=cy,
                            has_distortion=False, _type="PERSPECTIVE")
        return cam

    return types.SimpleNamespace(make_camera=make_camera, make_rays_worldspace=ns["make_rays_worldspace"],
                                 make_near_far_from_bound=ns["make_near_far_from_bound"], train_rays=train_rays, array=_j)


# ----------------------------------------------------------------------------------------------------------------------
# The NeRF model (models/nerfs.py:27-128 NeRF / CoordinateBasedMLP, :216-238 trunc_exp, :422-454 make_nerf_ngp) and the
# spherical-harmonics direction encoder (models/encoders.py:365-406), unmodified.  flax's Dense (x @ kernel, no bias
# here) and sigmoid are library code and are restated; layer order, widths, splits, the concatenation order of
# [x | SH(dir) | appearance], the activations and trunc_exp's custom backward rule are the reference's.
_module_stack = []


def _compact(fn):
    """nn.compact: tracks the module being applied so that submodules created inside find their parameters by flax's
    auto-naming (Dense_0, Dense_1, ... in creation order)."""

    @functools.wraps(fn)
    def wrapper(self, *a, **k):
        _module_stack.append(self)
        object.__setattr__(self, "_auto_index", {})
        try:
            return fn(self, *a, **k)
        finally:
            _module_stack.pop()

    return wrapper


class Dense:
    """flax.linen.Dense without bias: y = x @ kernel, kernel [in, out] bound on the parent as ``Dense_<i>``."""

    def __init__(self, features, use_bias=True, kernel_init=None, **kw):
        assert not use_bias, "the reference's MLPs are bias-free (models/nerfs.py:115-127)"
        self.features = features

    def __call__(self, x):
        parent = _module_stack[-1]
        i = parent._auto_index.get("Dense", 0)
        parent._auto_index["Dense"] = i + 1
        kernel = parent._bound_params[f"Dense_{i}"]
        assert kernel.shape == (x.shape[-1], self.features) and kernel.dtype == np.float32, (kernel.shape, x.shape, self.features)
        return _j(np.matmul(np.asarray(x, np.float32), kernel))


class CustomVjp:
    """jax.custom_vjp: the primal function, with the fwd/bwd rules kept for inspection (``.bwd(aux, g)``)."""

    def __init__(self, fn):
        self.fn, self.fwd, self.bwd = fn, None, None
        functools.update_wrapper(self, fn)

    def __call__(self, *a, **k):
        return self.fn(*a, **k)

    def defvjp(self, fwd, bwd):
        self.fwd, self.bwd = fwd, bwd


def install_nerf():
    """Imports the reference's models/encoders.py and models/nerfs.py unmodified; returns the ``models.nerfs`` module."""
    import typing
    install()  # jax / flax / chex / utils stand-ins + models.encoders
    jax = sys.modules["jax"]
    jnp = _JnpForTypes("jax.numpy")  # results carry .at (the SH encoder fills its output with .at[..., i].set)
    jax.numpy = jnp
    sys.modules["jax.numpy"] = jnp
    jax.custom_vjp = CustomVjp
    nn_mod = sys.modules["flax.linen"]
    nn_mod.compact, nn_mod.Dense = _compact, Dense
    nn_mod.relu = lambda x: _j(np.maximum(np.asarray(x), np.float32(0)))
    nn_mod.sigmoid = lambda x: _j((np.float32(1) / (np.float32(1) + np.exp(-np.asarray(x, np.float32)))).astype(np.float32))
    nn_mod.initializers = types.SimpleNamespace(glorot_uniform=lambda: None)
    init_mod = _Stub("jax.nn.initializers")
    init_mod.Initializer = object
    jax.nn = _Stub("jax.nn")
    jax.nn.initializers = init_mod
    sys.modules.update({"jax.nn": jax.nn, "jax.nn.initializers": init_mod})
    chex = sys.modules["chex"]
    chex.assert_axis_dimension = lambda *a, **k: None
    sys.modules["utils.common"].mkValueError = lambda **kw: ValueError(str(kw))
    rt = sys.modules["utils.types"]
    rt.ActivationType = rt.DirectionalEncodingType = rt.PositionalEncodingType = typing.Any
    # models.encoders must see the .at-capable jnp as well: re-import it under its package name
    models = types.ModuleType("models")
    models.__path__ = []
    sys.modules["models"] = models
    for name in ("encoders", "nerfs"):
        spec = importlib.util.spec_from_file_location(f"models.{name}", os.path.join(REFERENCE, "models", f"{name}.py"))
        module = importlib.util.module_from_spec(spec)
        sys.modules[f"models.{name}"] = module
        spec.loader.exec_module(module)
        setattr(models, name, module)
    return sys.modules["models.nerfs"]


# ----------------------------------------------------------------------------------------------------------------------
# Loss and optimizer of the training step (app/nerf/_utils.py:19-77 make_optimizer, :117-162 loss_fn nested in
# train_step; utils/data.py:443-464 blend_rgba_image_array), unmodified.  optax is a third-party dependency that is not
# on disk (the reference's flake pins nixpkgs' python3Packages.optax): the few primitives the reference configures are
# restated from their published definitions -- huber_loss, exponential_decay, scale_by_adam with eps_root, scaling by
# -learning_rate, add_decayed_weights, multi_transform / chain over a pytree of dicts.
def _tree_map(fn, *trees):
    t0 = trees[0]
    if isinstance(t0, dict):
        return {k: _tree_map(fn, *[t[k] for t in trees]) for k in t0}
    return fn(*trees)


def _mask_lookup(mask, path_value_tree):
    """optax masks may be prefixes of the parameter tree: a bool at an inner node covers the whole subtree."""
    if isinstance(path_value_tree, dict):
        return {k: _mask_lookup(mask[k] if isinstance(mask, dict) else mask, v) for k, v in path_value_tree.items()}
    return bool(mask)


class OptaxShim(types.ModuleType):
    GradientTransformation = object

    def __init__(self):
        super().__init__("optax")
        self.calls = []  # what the reference configured, in call order

    @staticmethod
    def huber_loss(predictions, targets=None, delta=1.0):
        err = np.asarray(predictions, np.float32) - np.asarray(targets, np.float32)
        abs_err = np.abs(err)
        quad = np.minimum(abs_err, np.float32(delta))
        return _j((np.float32(0.5) * quad * quad + np.float32(delta) * (abs_err - quad)).astype(np.float32))

    def exponential_decay(self, init_value, transition_steps, decay_rate, transition_begin=0, staircase=False, end_value=None):
        self.calls.append(("exponential_decay", dict(init_value=init_value, transition_steps=transition_steps, decay_rate=decay_rate,
                                                     transition_begin=transition_begin, staircase=staircase, end_value=end_value)))

        def schedule(count):
            decreased = count - transition_begin
            p = decreased / transition_steps
            if staircase:
                p = np.floor(p)
            value = init_value * decay_rate ** p if decreased > 0 else init_value
            if end_value is not None:
                value = max(value, end_value) if decay_rate < 1 else min(value, end_value)
            return value

        self.last_schedule = schedule
        return schedule

    def adam(self, learning_rate, b1=0.9, b2=0.999, eps=1e-8, eps_root=0.0):
        self.calls.append(("adam", dict(learning_rate=learning_rate if not callable(learning_rate) else "schedule", b1=b1, b2=b2,
                                        eps=eps, eps_root=eps_root)))

        def init(params):
            return dict(count=0, mu=_tree_map(np.zeros_like, params), nu=_tree_map(np.zeros_like, params))

        def update(grads, state, params=None):
            count = state["count"] + 1
            mu = _tree_map(lambda m, g: b1 * m + (1 - b1) * g, state["mu"], grads)
            nu = _tree_map(lambda v, g: b2 * v + (1 - b2) * g * g, state["nu"], grads)
            lr = learning_rate(state["count"]) if callable(learning_rate) else learning_rate
            upd = _tree_map(lambda m, v: (-lr * (m / (1 - b1 ** count)) / (np.sqrt(v / (1 - b2 ** count) + eps_root) + eps)).astype(np.float32),
                            mu, nu)
            return upd, dict(count=count, mu=mu, nu=nu)

        return types.SimpleNamespace(init=init, update=update)

    def multi_transform(self, transforms, param_labels):
        self.calls.append(("multi_transform", dict(labels=param_labels, transforms=sorted(transforms))))

        def init(params):
            return {k: transforms[param_labels[k]].init(v) for k, v in params.items()}

        def update(grads, state, params=None):
            out, new_state = {}, {}
            for k in grads:
                out[k], new_state[k] = transforms[param_labels[k]].update(grads[k], state[k], None if params is None else params[k])
            return out, new_state

        return types.SimpleNamespace(init=init, update=update)

    def add_decayed_weights(self, weight_decay=0.0, mask=None):
        self.calls.append(("add_decayed_weights", dict(weight_decay=weight_decay, mask=mask)))

        def update(updates, state, params):
            m = _mask_lookup(mask, params)
            return _tree_map(lambda u, p, on: (u + np.float32(weight_decay) * p).astype(np.float32) if on else u, updates, params, m), state

        return types.SimpleNamespace(init=lambda params: None, update=update)

    def chain(self, *transforms):
        self.calls.append(("chain", dict(n=len(transforms))))

        def init(params):
            return [t.init(params) for t in transforms]

        def update(grads, state, params=None):
            new_state = []
            for t, s in zip(transforms, state):
                grads, s = t.update(grads, s, params)
                new_state.append(s)
            return grads, new_state

        return types.SimpleNamespace(init=init, update=update)


def install_loss_and_optimizer(scripted_random):
    """Returns a namespace with the reference's ``make_optimizer``, ``blend_rgba_image_array`` and a callable
    ``loss(pred_rgbds, ray_is_valid, gt_rgba_f32)`` that runs the reference's nested ``loss_fn`` with the renderer
    replaced by its outputs."""
    import typing
    jnp = _JnpForTypes("jax.numpy")
    jax = _Stub("jax")
    jax.Array, jax.numpy, jax.random = np.ndarray, jnp, scripted_random
    jax.tree_util = types.SimpleNamespace(tree_reduce=lambda fn, tree: functools.reduce(fn, list(tree.values())))
    optax = OptaxShim()
    chex = _Stub("chex")
    for name in ("assert_shape", "assert_type"):
        setattr(chex, name, lambda *a, **k: None)
    pil = types.SimpleNamespace(Image=type("Image", (), {}))
    data_ns = dict(jnp=jnp, jax=jax, np=np, chex=chex, Image=pil, Tuple=typing.Tuple)
    exec(compile(_extract_functions(os.path.join(REFERENCE, "utils", "data.py"), {"blend_rgba_image_array"}), "utils/data.py", "exec"), data_ns)
    utils_path = os.path.join(REFERENCE, "app", "nerf", "_utils.py")
    ns = dict(optax=optax)
    exec(compile(_extract_functions(utils_path, {"make_optimizer"}), "app/nerf/_utils.py", "exec"), ns)
    tree = ast.parse(open(utils_path).read())
    train_step = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "train_step")
    loss_fn = next(n for n in train_step.body if isinstance(n, ast.FunctionDef) and n.name == "loss_fn")
    train_ns = dict(jnp=jnp, jax=jax, jran=scripted_random, optax=optax,
                    data=types.SimpleNamespace(blend_rgba_image_array=data_ns["blend_rgba_image_array"]))
    exec(compile(ast.Module(body=[loss_fn], type_ignores=[]), "app/nerf/_utils.py", "exec"), train_ns)

    def loss(pred_rgbds, ray_is_valid, gt_rgba_f32):
        n = pred_rgbds.shape[0]
        valid = _j(np.asarray(ray_is_valid, bool))
        train_ns.update(
            make_rays_worldspace=lambda: (_j(np.zeros((n, 3), np.float32)), _j(np.zeros((n, 3), np.float32))),
            view_idcs=None, total_samples=0,
            state=types.SimpleNamespace(use_background_model=False, render=types.SimpleNamespace(random_bg=True),
                                        replace=lambda **kw: None),
            render_rays_train=lambda **kw: (dict(ray_is_valid=valid, n_valid_rays=valid.sum()), _j(np.asarray(pred_rgbds, np.float32)), 0))
        value, metrics = train_ns["loss_fn"]({}, _j(np.asarray(gt_rgba_f32, np.float32)), None)
        return np.float32(value), metrics

    return types.SimpleNamespace(make_optimizer=ns["make_optimizer"], optax=optax, loss=loss,
                                 blend_rgba_image_array=data_ns["blend_rgba_image_array"], array=_j)


def install_renderer(oracle_module):
    """The reference's inference renderer, unmodified: ``render_image_inference`` and ``march_and_integrate_inference``
    (models/renderers/cuda.py:165-373), the Python wrappers ``march_rays_inference`` / ``integrate_rays_inference`` of
    volume-rendering-jax (marching/__init__.py:96-157, integrating/__init__.py:62-110) and ``f32_to_u8``
    (utils/data.py:42-43).  The two custom-call primitives the wrappers bind are served by the C oracle.  Returns a
    namespace with ``render_image_inference`` and the pieces a caller needs to build the ``state`` it reads."""
    import math
    import typing
    rays = install_rays()
    jnp = _JnpForTypes("jax.numpy")
    jax = _Stub("jax")
    jax.Array, jax.numpy = np.ndarray, jnp
    jax.jit = lambda fun=None, **kw: fun if fun is not None else (lambda f: f)
    jax.lax = types.SimpleNamespace(stop_gradient=lambda x: x)

    class _Prim:
        def __init__(self, fn):
            self.bind = fn

    def march_bind(rays_o, rays_d, t_starts, t_ends, bitfield, next_ray_index_in, terminated, indices, **static):
        nri, idx, ns, _, xyzs, dss, zs, tso = oracle_module.march_rays_inference(
            static["diagonal_n_steps"], static["K"], static["G"], static["march_steps_cap"], static["bound"],
            static["stepsize_portion"], np.asarray(rays_o), np.asarray(rays_d), np.asarray(t_starts), np.asarray(t_ends),
            np.asarray(bitfield), np.asarray(next_ray_index_in), np.asarray(terminated), np.asarray(indices))
        return tuple(_j(x) for x in (nri, idx, ns, tso, xyzs, dss, zs))

    def integrate_bind(rays_bg, rays_rgbd, rays_T, n_samples, indices, dss, z_vals, drgbs):
        cnt, term, rgbd_o, T_o = oracle_module.integrate_rays_inference(
            np.asarray(rays_bg), np.asarray(rays_rgbd), np.asarray(rays_T), np.asarray(n_samples), np.asarray(indices),
            np.asarray(dss), np.asarray(z_vals), np.asarray(drgbs, np.float32), raw=True)
        return _j(cnt), _j(term), _j(rgbd_o), _j(T_o)

    impl = types.SimpleNamespace(march_rays_inference_p=_Prim(march_bind), integrate_rays_inference_p=_Prim(integrate_bind))

    def load_wrapper(rel, name):
        path = os.path.join(REFERENCE, "deps", "volume-rendering-jax", "src", "volrendjax", rel, "__init__.py")
        ns = dict(jax=jax, jnp=jnp, Tuple=typing.Tuple, impl=impl)
        exec(compile(_extract_functions(path, {name}), f"volrendjax/{rel}/__init__.py", "exec"), ns)
        return ns[name]

    data_ns = dict(jnp=jnp, jax=jax)
    exec(compile(_extract_functions(os.path.join(REFERENCE, "utils", "data.py"), {"f32_to_u8"}), "utils/data.py", "exec"), data_ns)
    common_ns = dict(functools=functools, jax=jax, Any=typing.Any, Hashable=typing.Hashable, Sequence=typing.Sequence,
                     Iterable=typing.Iterable, xc=_Stub("xc"))
    exec(compile(_extract_functions(os.path.join(REFERENCE, "utils", "common.py"), {"jit_jaxfn_with"}), "utils/common.py", "exec"), common_ns)
    path = os.path.join(REFERENCE, "models", "renderers", "cuda.py")
    tree = ast.parse(open(path).read())
    wanted = {"MarchAndIntegrateInferencePayload", "march_and_integrate_inference", "render_image_inference"}
    body = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in wanted]
    assert {n.name for n in body} == wanted
    placeholder = type("Placeholder", (), {})
    ns = dict(jax=jax, jnp=jnp, jran=_Stub("jax.random"), math=math, dataclass=dataclasses.dataclass, Callable=typing.Callable,
              jit_jaxfn_with=common_ns["jit_jaxfn_with"], FrozenVariableDict=typing.Any, RigidTransformation=typing.Any,
              NeRFState=typing.Any, Camera=placeholder, CameraOverrideOptions=type("CameraOverrideOptions", (), {}),
              march_rays_inference=load_wrapper("marching", "march_rays_inference"),
              integrate_rays_inference=load_wrapper("integrating", "integrate_rays_inference"),
              make_rays_worldspace=rays.make_rays_worldspace, make_near_far_from_bound=rays.make_near_far_from_bound,
              f32_to_u8=data_ns["f32_to_u8"])
    exec(compile(ast.Module(body=body, type_ignores=[]), "models/renderers/cuda.py", "exec"), ns)
    return types.SimpleNamespace(render_image_inference=ns["render_image_inference"], make_camera=rays.make_camera, array=_j)


def install_train_forward(oracle_module, scripted_random):
    """The whole forward of the reference's training step from its own source: the ray construction and ``loss_fn``
    nested in ``train_step`` (app/nerf/_utils.py:93-162), ``render_rays_train`` (models/renderers/cuda.py:100-162), the
    ``march_rays`` / ``integrate_rays`` wrappers and the primal of ``__integrate_rays`` (volume-rendering-jax), the
    ``make_nerf_ngp`` model with the pure-JAX hash-grid encoder, ``blend_rgba_image_array``.  Only the two CUDA
    primitives (served by the C oracle) and the library calls (flax Dense / sigmoid, optax.huber_loss) are stand-ins.
    Returns ``forward(perm, transforms, camera, table, weights, bitfield, rgbas_u8, total_samples) -> (loss, metrics,
    pred_rgbds)``."""
    import typing
    nerfs = install_nerf()
    rays = install_rays()
    lo = install_loss_and_optimizer(scripted_random)
    jnp = _JnpForTypes("jax.numpy")
    jax = _Stub("jax")
    jax.Array, jax.numpy, jax.random, jax.custom_vjp = np.ndarray, jnp, scripted_random, CustomVjp
    jax.jit = lambda fun=None, **kw: fun if fun is not None else (lambda f: f)
    jax.tree_util = types.SimpleNamespace(tree_reduce=lambda fn, tree: functools.reduce(fn, list(tree.values())))
    vr = os.path.join(REFERENCE, "deps", "volume-rendering-jax", "src", "volrendjax")

    class _Prim:
        def __init__(self, fn):
            self.bind = fn

    def march_bind(rays_o, rays_d, t_starts, t_ends, noises, bitfield, **st):
        out = oracle_module.march_rays(st["total_samples"], st["diagonal_n_steps"], st["K"], st["G"], st["bound"],
                                       st["stepsize_portion"], np.asarray(rays_o), np.asarray(rays_d), np.asarray(t_starts),
                                       np.asarray(t_ends), np.asarray(noises), np.asarray(bitfield), raw=True)
        return tuple(_j(np.asarray(x)) for x in out)

    def integrate_bind(start, ns, bgs, dss, z_vals, drgbs):
        mbs, rgbd, opac = oracle_module.integrate_rays(0.0, np.asarray(start), np.asarray(ns), np.asarray(bgs), np.asarray(dss),
                                                       np.asarray(z_vals), np.asarray(drgbs, np.float32))
        return _j(np.array([mbs], np.uint32)), _j(rgbd), _j(opac)

    impl_int_ns = dict(jax=jax, Tuple=typing.Tuple, integrate_rays_p=_Prim(integrate_bind))
    exec(compile(_extract_functions(os.path.join(vr, "integrating", "impl.py"), {"__integrate_rays"}), "integrating/impl.py", "exec"), impl_int_ns)
    impl_int = types.SimpleNamespace(**{"__integrate_rays": impl_int_ns["__integrate_rays"]})
    int_ns = dict(jax=jax, Tuple=typing.Tuple, impl=impl_int)
    exec(compile(_extract_functions(os.path.join(vr, "integrating", "__init__.py"), {"integrate_rays"}), "integrating/__init__.py", "exec"), int_ns)
    march_ns = dict(jax=jax, jnp=jnp, Tuple=typing.Tuple, impl=types.SimpleNamespace(march_rays_p=_Prim(march_bind)))
    exec(compile(_extract_functions(os.path.join(vr, "marching", "__init__.py"), {"march_rays"}), "marching/__init__.py", "exec"), march_ns)
    common_ns = dict(functools=functools, jax=jax, Any=typing.Any, Hashable=typing.Hashable, Sequence=typing.Sequence,
                     Iterable=typing.Iterable, xc=_Stub("xc"))
    exec(compile(_extract_functions(os.path.join(REFERENCE, "utils", "common.py"), {"jit_jaxfn_with"}), "utils/common.py", "exec"), common_ns)
    cuda_ns = dict(jax=jax, jnp=jnp, jran=scripted_random, jit_jaxfn_with=common_ns["jit_jaxfn_with"], NeRFState=typing.Any,
                   make_near_far_from_bound=rays.make_near_far_from_bound, march_rays=march_ns["march_rays"],
                   integrate_rays=int_ns["integrate_rays"])
    exec(compile(_extract_functions(os.path.join(REFERENCE, "models", "renderers", "cuda.py"), {"render_rays_train"}),
                 "models/renderers/cuda.py", "exec"), cuda_ns)
    utils_path = os.path.join(REFERENCE, "app", "nerf", "_utils.py")
    tree = ast.parse(open(utils_path).read())
    train_step = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "train_step")
    nested = [n for n in train_step.body if isinstance(n, ast.FunctionDef) and n.name in ("make_rays_worldspace", "loss_fn")]
    assert len(nested) == 2
    train_ns = dict(jnp=jnp, jax=jax, jran=scripted_random, optax=lo.optax, Tuple=typing.Tuple,
                    data=types.SimpleNamespace(blend_rgba_image_array=lo.blend_rgba_image_array),
                    render_rays_train=cuda_ns["render_rays_train"])
    exec(compile(ast.Module(body=nested, type_ignores=[]), "app/nerf/_utils.py", "exec"), train_ns)
    model = nerfs.make_nerf_ngp(bound=1.0, inference=False)

    def forward(perm, transforms, camera, table, weights, bitfield, rgbas_u8, total_samples, n_views):
        model.position_encoder.bind_params(**{"latent codes stored on grid vertices": table})
        model.density_mlp.bind_params(Dense_0=weights["density_w0"], Dense_1=weights["density_w1"])
        model.rgb_mlp.bind_params(Dense_0=weights["rgb_w0"], Dense_1=weights["rgb_w1"], Dense_2=weights["rgb_w2"])
        perm = _j(np.asarray(perm, np.uint32))

        class State:
            use_background_model = False
            render = types.SimpleNamespace(random_bg=True)
            raymarch = types.SimpleNamespace(perturb=True, diagonal_n_steps=1024, density_grid_res=128)
            scene_meta = types.SimpleNamespace(bound=1.0, cascades=1, stepsize_portion=0.0, camera=camera)
            ogrid = types.SimpleNamespace(occupancy=_j(np.asarray(bitfield)))
            nerf_fn = staticmethod(lambda variables, xyzs, dirs, app: model(xyzs, dirs, app))
            params = None

            def replace(self, **kw):
                new = State()
                new.__dict__.update(self.__dict__)
                new.__dict__.update(kw)
                return new

        train_ns.update(state=State(), total_samples=total_samples,
                        scene=types.SimpleNamespace(meta=types.SimpleNamespace(camera=camera), transforms=_j(np.asarray(transforms, np.float32))),
                        view_idcs=jnp.floor_divide(perm, camera.n_pixels), pixel_idcs=jnp.mod(perm, camera.n_pixels))
        params = {"nerf": None, "appearance_embeddings": _j(np.zeros((n_views, 0), np.float32))}
        gt = _j((np.asarray(rgbas_u8[np.asarray(perm)]).astype(np.float32) / np.float32(255)))  # _utils.py:165
        loss, metrics = train_ns["loss_fn"](params, gt, None)
        return np.float32(loss), metrics

    return types.SimpleNamespace(forward=forward, make_camera=rays.make_camera)


def install_cadence():
    """``NeRFState.update_ogrid_interval`` / ``should_call_update_ogrid`` / ``should_update_all_ogrid_cells``
    (utils/types.py:1380-1396), unmodified; returns ``state(step)``."""
    path = os.path.join(REFERENCE, "utils", "types.py")
    tree = ast.parse(open(path).read())
    state_cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "NeRFState")
    wanted = {"update_ogrid_interval", "should_call_update_ogrid", "should_update_all_ogrid_cells"}
    methods = [n for n in state_cls.body if isinstance(n, ast.FunctionDef) and n.name in wanted]
    assert {m.name for m in methods} == wanted
    cls = ast.ClassDef(name="Cadence", bases=[], keywords=[], body=methods, decorator_list=[], type_params=[])
    module = ast.Module(body=[cls], type_ignores=[])
    ast.fix_missing_locations(module)
    ns = {}
    exec(compile(module, "utils/types.py", "exec"), ns)

    def state(step):
        obj = ns["Cadence"]()
        obj.step = step
        return obj

    return state


# ----------------------------------------------------------------------------------------------------------------------
# Abstract-evaluation contracts of the ops (volume-rendering-jax: {marching,integrating,packbits,morton3d}/abstract.py):
# output shapes / dtypes and the errors raised for malformed operands, unmodified.  chex's assertions are restated
# (they raise AssertionError); jax.dtypes.canonicalize_dtype maps 64-bit types to 32-bit (x64 is off in the reference).
class ShapedArray:
    def __init__(self, shape, dtype):
        self.shape, self.dtype = tuple(shape), np.dtype(dtype)

    @property
    def ndim(self):
        return len(self.shape)


def _chex_for_contracts():
    chex = types.ModuleType("chex")

    def _each(x):
        return list(x) if isinstance(x, (list, tuple)) else [x]

    def assert_shape(x, shape):
        for a in _each(x):
            if tuple(a.shape) != tuple(shape):
                raise AssertionError(f"shape {a.shape} != {tuple(shape)}")

    def assert_type(x, dtype):
        for a in _each(x):
            if np.dtype(a.dtype) != np.dtype(dtype):
                raise AssertionError(f"dtype {a.dtype} != {np.dtype(dtype)}")

    def assert_rank(x, rank):
        for a in _each(x):
            if len(a.shape) != rank:
                raise AssertionError(f"rank {len(a.shape)} != {rank}")

    def assert_equal_shape(xs):
        if len({tuple(a.shape) for a in xs}) > 1:
            raise AssertionError("shapes differ")

    def assert_axis_dimension(x, axis, expected):
        if x.shape[axis] != expected:
            raise AssertionError(f"axis {axis} has {x.shape[axis]} != {expected}")

    def assert_scalar_positive(v):
        if not v > 0:
            raise AssertionError(f"{v} is not positive")

    def assert_scalar_non_negative(v):
        if not v >= 0:
            raise AssertionError(f"{v} is negative")

    def assert_scalar(v):
        if not isinstance(v, (int, float)):
            raise AssertionError(f"{v!r} is not a scalar")

    array_assert_type = assert_type

    def assert_type(x, dtype):  # chex accepts Python scalars too: their type must be (a subtype of) the expected one
        if all(isinstance(a, (int, float)) for a in _each(x)):
            for a in _each(x):
                if not (isinstance(a, dtype) and not (dtype is float and isinstance(a, bool))) or (dtype is float and isinstance(a, int)):
                    raise AssertionError(f"{a!r} is not of type {dtype}")
            return
        array_assert_type(x, dtype)

    for f in (assert_shape, assert_type, assert_rank, assert_equal_shape, assert_axis_dimension, assert_scalar_positive,
              assert_scalar_non_negative, assert_scalar):
        setattr(chex, f.__name__, f)
    return chex


def install_abstract():
    """Returns {name: function} with every ``*_abstract`` rule of volume-rendering-jax, and the ShapedArray class."""
    def canonicalize(dt):
        dt = np.dtype(dt)
        return {np.dtype(np.float64): np.dtype(np.float32), np.dtype(np.int64): np.dtype(np.int32),
                np.dtype(np.uint64): np.dtype(np.uint32)}.get(dt, dt)

    jnp = _numpy_namespace("jax.numpy")
    jax = _Stub("jax")
    jax.numpy, jax.ShapedArray = jnp, ShapedArray
    jax.dtypes = types.SimpleNamespace(canonicalize_dtype=canonicalize)
    saved = {k: sys.modules.get(k) for k in ("jax", "jax.numpy", "chex")}
    sys.modules.update({"jax": jax, "jax.numpy": jnp, "chex": _chex_for_contracts()})
    rules = {}
    try:
        paths = {pkg: os.path.join(REFERENCE, "deps", "volume-rendering-jax", "src", "volrendjax", pkg, "abstract.py")
                 for pkg in ("marching", "integrating", "packbits", "morton3d")}
        paths["hashgrid_tcnn"] = os.path.join(REFERENCE, "deps", "jax-tcnn", "src", "jaxtcnn", "hashgrid_tcnn", "abstract.py")
        for pkg, path in paths.items():
            spec = importlib.util.spec_from_file_location(f"reference_volrendjax_{pkg}_abstract", path)
            module = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(module)
            rules.update({k: v for k, v in vars(module).items() if k.endswith("_abstract")})
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return rules, ShapedArray

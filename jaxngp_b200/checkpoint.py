"""Array interchange with the reference's checkpoints: the flat buffers of this package
<-> the parameter tree of the reference's ``NeRFState`` (``app/nerf/train.py:172-199``), so that a model trained there
renders through these kernels and the other way round.

Names and shapes follow the reference's own modules (checked against its unmodified code in
the CPU test ``test_checkpoint_tree_binds_to_the_reference_model``):

    step                                                           int      TrainState.step
    params/nerf/position_encoder/latent codes stored on grid vertices  f32[rows, F]   models/encoders.py:105-114
    params/nerf/density_mlp/Dense_{0,1}/kernel                     f32[32,64] [64,16]   models/nerfs.py:89-128,422-454
    params/nerf/rgb_mlp/Dense_{0,1,2}/kernel                       f32[32,64] [64,64] [64,3]
    params/bg                                                      None     (scene_meta.bg is false for NeRF-synthetic)
    params/appearance_embeddings                                   f32[n_frames, 0]     app/nerf/train.py:190-196
    ogrid/{density, occ_mask, occupancy, alive_indices}            f32 / bool / u8 / u32     utils/types.py:93-144

Two containers: ``.npz`` with ``/``-joined keys, and the msgpack wire format of ``flax.serialization`` that
``flax.training.checkpoints`` writes (``checkpoint_<step>`` files; ``app/nerf/train.py:200``, ``app/nerf/test.py:53-63``).
flax is a third-party dependency that is not on disk: the container is restated from its published format (ndarray =
ExtType 1 holding msgpack((shape, dtype name, C-order bytes)), numpy scalar = ExtType 3, tuples/lists as dicts keyed by
their index) and is therefore unpinned; the tree that goes into it is pinned.  The optimizer state is exchanged as
plain ``adam_m`` / ``adam_v`` trees shaped like ``params/nerf`` under the private key ``opt_state_b200`` (npz only):
optax's own ``opt_state`` layout depends on its version and is not restated.

Host-side I/O only: nothing here computes on the path.
"""
import os

import numpy as np

from . import nerf as nerf_mod

TABLE_NAME = "latent codes stored on grid vertices"  # models/encoders.py:106
_MLP_TREE = (("density_mlp", "Dense_0", "density_w0"), ("density_mlp", "Dense_1", "density_w1"),
             ("rgb_mlp", "Dense_0", "rgb_w0"), ("rgb_mlp", "Dense_1", "rgb_w1"), ("rgb_mlp", "Dense_2", "rgb_w2"))
_SHAPES = {name: (i, o) for name, i, o in nerf_mod.MLP_SHAPES}


class CheckpointError(ValueError):
    pass


def _np(a):
    if hasattr(a, "detach"):  # torch tensor (any device)
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a)


# ---------------------------------------------------------------------------------------------- parameter tree
def nerf_param_tree(table, mlp_flat) -> dict:
    """Flat buffers -> ``params["nerf"]`` of the reference (numpy arrays, copies)."""
    table, flat = _np(table).astype(np.float32, copy=True), _np(mlp_flat).astype(np.float32, copy=False).reshape(-1)
    if table.ndim != 2:
        raise CheckpointError(f"hash table must be [rows, F], got {table.shape}")
    if flat.size < nerf_mod.MLP_NUMEL:
        raise CheckpointError(f"MLP buffer holds {flat.size} weights, expected {nerf_mod.MLP_NUMEL}")
    tree = {"position_encoder": {TABLE_NAME: table}, "density_mlp": {}, "rgb_mlp": {}}
    off = 0
    by_name = {}
    for name, i, o in nerf_mod.MLP_SHAPES:  # kernels are [in, out], row-major: flax's Dense layout
        by_name[name] = flat[off:off + i * o].reshape(i, o).copy()
        off += i * o
    for module, layer, name in _MLP_TREE:
        tree[module][layer] = {"kernel": by_name[name]}
    return tree


def flat_from_nerf_param_tree(tree: dict, rows: int = None, F: int = None):
    """``params["nerf"]`` of the reference -> (table f32[rows, F], mlp_flat f32[MLP_NUMEL]).  Raises CheckpointError for a
    missing module, a wrong shape or a non-float32 leaf instead of guessing."""
    try:
        table = np.asarray(tree["position_encoder"][TABLE_NAME])
    except (KeyError, TypeError) as exc:
        raise CheckpointError(f"params/nerf/position_encoder/{TABLE_NAME!r} is missing: this is not a hash-grid NGP model") from exc
    if table.dtype != np.float32 or table.ndim != 2:
        raise CheckpointError(f"hash table must be float32 [rows, F], got {table.dtype} {table.shape}")
    if (rows is not None and table.shape[0] != rows) or (F is not None and table.shape[1] != F):
        raise CheckpointError(
            f"hash table is {table.shape}, this model expects ({rows}, {F}).  Tables written by the training encoder have "
            "8-aligned level sizes (models/encoders.py:96), the tiny-cuda-nn layout of the inference model does not (:275): "
            "build the model with inference=False to load a checkpoint trained with the pure-JAX encoder")
    parts = {}
    for module, layer, name in _MLP_TREE:
        try:
            k = np.asarray(tree[module][layer]["kernel"])
        except (KeyError, TypeError) as exc:
            raise CheckpointError(f"params/nerf/{module}/{layer}/kernel is missing") from exc
        if k.dtype != np.float32 or k.shape != _SHAPES[name]:
            raise CheckpointError(f"params/nerf/{module}/{layer}/kernel is {k.dtype} {k.shape}, expected float32 {_SHAPES[name]} "
                                  "(appearance embeddings and other widths are not supported)")
        parts[name] = k
    extra = {m: set(tree[m]) - {l for mm, l, _ in _MLP_TREE if mm == m} for m in ("density_mlp", "rgb_mlp")}
    if any(extra.values()):
        raise CheckpointError(f"unexpected layers {extra}: deeper MLPs are not supported")
    flat = np.concatenate([parts[name].reshape(-1) for name, _, _ in nerf_mod.MLP_SHAPES]).astype(np.float32)
    return np.ascontiguousarray(table), flat


def make_state(step, table, mlp_flat, grid=None, n_frames=0, adam_m=None, adam_v=None, table_numel=None) -> dict:
    """The pytree-node fields of the reference's ``NeRFState`` as nested dicts of numpy arrays.  ``grid``: an object with
    ``density`` / ``occ_mask`` / ``occupancy`` (``ogrid.OccupancyDensityGrid``).  ``adam_m`` / ``adam_v``: flat moment
    buffers laid out like [table | MLP weights] (the trainer's, world_size 1)."""
    state = {"step": int(step),
             "params": {"nerf": nerf_param_tree(table, mlp_flat), "bg": None,
                        "appearance_embeddings": np.zeros((int(n_frames), 0), np.float32)},
             # flax's from_state_dict wants every pytree field of TrainState present; inference (app/nerf/test.py) never
             # reads this one, and optax's own layout is not restated (module docstring)
             "opt_state": None}
    if grid is not None:
        density = _np(grid.density).astype(np.float32, copy=False)
        state["ogrid"] = {"density": density, "occ_mask": _np(grid.occ_mask).astype(np.bool_, copy=False),
                          "occupancy": _np(grid.occupancy).astype(np.uint8, copy=False),
                          # types.py:139; the reference re-marks after a load anyway (train.py:206)
                          "alive_indices": (np.arange(density.shape[0], dtype=np.uint32) if getattr(grid, "alive_indices", None) is None
                                            else _np(grid.alive_indices).astype(np.int64).astype(np.uint32))}
        if state["ogrid"]["occupancy"].shape[0] * 8 != density.shape[0] or state["ogrid"]["occ_mask"].shape != density.shape:
            raise CheckpointError("density grid arrays disagree in size")
    if adam_m is not None and adam_v is not None:
        n_table = _np(table).size if table_numel is None else int(table_numel)
        state["opt_state_b200"] = {
            key: nerf_param_tree(_np(buf)[:n_table].reshape(_np(table).shape), _np(buf)[n_table:n_table + nerf_mod.MLP_NUMEL])
            for key, buf in (("adam_m", adam_m), ("adam_v", adam_v))}
    return state


def find_adam_moments(opt_state):
    """Best effort, for checkpoints WRITTEN BY THE REFERENCE: optax's ``ScaleByAdamState`` serialises as a dict with
    ``count`` / ``mu`` / ``nu`` wherever the surrounding ``chain`` / ``multi_transform`` / ``masked`` wrappers of the
    installed optax version put it; the network optimizer's is the one whose ``mu`` carries a ``nerf`` sub-tree
    (app/nerf/_utils.py:46-56).  Returns (mu["nerf"], nu["nerf"], count) or None."""
    if not isinstance(opt_state, dict):
        return None
    mu, nu = opt_state.get("mu"), opt_state.get("nu")
    if isinstance(mu, dict) and isinstance(nu, dict) and isinstance(mu.get("nerf"), dict) and isinstance(nu.get("nerf"), dict) \
            and "position_encoder" in mu["nerf"]:
        return mu["nerf"], nu["nerf"], opt_state.get("count")
    for v in opt_state.values():
        found = find_adam_moments(v)
        if found is not None:
            return found
    return None


def state_from_trainer(trainer, n_frames=None) -> dict:
    """Snapshot of a (world_size 1) ``trainer.Trainer``."""
    if trainer.world_size != 1:
        raise CheckpointError("optimizer moments are sharded across ranks: snapshot a world_size-1 trainer, or gather first")
    n_frames = trainer.scene.n_views if n_frames is None else n_frames
    return make_state(int(trainer.step_dev.item()), trainer.table, trainer.mlp_flat, trainer.grid, n_frames,
                      trainer.adam_m, trainer.adam_v, trainer.table_numel)


def load_into_trainer(trainer, state: dict):
    """Copies parameters, density grid, step and (when present) the Adam moments of ``state`` into the trainer's
    device buffers.  Captured graphs keep working: buffers are overwritten in place.  Everything is validated before
    the first copy, so a refused state leaves the trainer untouched."""
    import torch

    table, flat = flat_from_nerf_param_tree(state["params"]["nerf"], *trainer.table.shape)
    if state["params"].get("bg") is not None:
        raise CheckpointError("background models are outside this path (scene_meta.bg)")
    ae = state["params"].get("appearance_embeddings")
    if ae is not None and np.asarray(ae).size:
        raise CheckpointError("appearance embeddings are outside this path (n_extra_learnable_dims > 0)")
    grid_arrays = {}
    if "ogrid" in state:
        for name in ("density", "occ_mask", "occupancy"):
            src = torch.from_numpy(np.ascontiguousarray(state["ogrid"][name]))
            dst = getattr(trainer.grid, name)
            if src.shape != dst.shape or src.dtype != dst.dtype:
                raise CheckpointError(f"ogrid/{name} is {src.dtype} {tuple(src.shape)}, this grid holds {dst.dtype} "
                                      f"{tuple(dst.shape)} (cascades / resolution differ)")
            grid_arrays[name] = src
    opt = state.get("opt_state_b200")
    if opt is None:  # a checkpoint written by the reference: look for the network optimizer's Adam moments
        found = find_adam_moments(state.get("opt_state"))
        if found is not None:
            opt = {"adam_m": found[0], "adam_v": found[1]}
    moments = {}
    if opt is not None:
        if trainer.world_size != 1:
            raise CheckpointError("optimizer moments can only be loaded into a world_size-1 trainer")
        for key in ("adam_m", "adam_v"):
            moments[key] = flat_from_nerf_param_tree(opt[key], *trainer.table.shape)

    dev = trainer.device
    if hasattr(trainer, "drop_prefetch"):
        trainer.drop_prefetch()  # a march prefetched against the old bitfield must not race with the copies below
    else:
        trainer._prefetched = None
    trainer.table.copy_(torch.from_numpy(table).to(dev))
    trainer.mlp_flat.copy_(torch.from_numpy(flat).to(dev))
    for name, src in grid_arrays.items():
        getattr(trainer.grid, name).copy_(src.to(dev))
    if "ogrid" in state:
        # the trainable-cell table (utils/types.py:1353-1358): without it the next full update would sample culled
        # cells again.  The reference's checkpoint stores the indices only; the per-cascade offsets are derived from them
        # (the reference re-marks after every restore, app/nerf/train.py:206 -- callers may still do that).
        alive = state["ogrid"].get("alive_indices")
        g = trainer.grid
        n_cells = int(g.density.shape[0])
        cascades = int(getattr(g, "K", 1))
        cells_per_cascade = n_cells // cascades
        if alive is None or np.asarray(alive).size in (0, n_cells):
            g.alive_indices = g.alive_indices_offset = None
        else:
            alive = np.sort(np.asarray(alive).astype(np.int64).reshape(-1))
            g.alive_indices = torch.from_numpy(alive.astype(np.int32)).to(dev)
            g.alive_indices_offset = [int(np.searchsorted(alive, c * cells_per_cascade)) for c in range(cascades + 1)]
    trainer.step = int(state["step"])
    trainer.step_dev.fill_(int(state["step"]))
    for key, (t, f) in moments.items():
        buf = getattr(trainer, key)
        buf[: trainer.table_numel].copy_(torch.from_numpy(t.reshape(-1)).to(dev))
        buf[trainer.table_numel:trainer.n_params].copy_(torch.from_numpy(f).to(dev))


def load_into_model(model, state: dict):
    """Copies the parameters of ``state`` into a ``nerf.NeRF`` (in place, on its device) and returns the occupancy
    bitfield u8[K*G^3/8] on that device, or None if the state carries no grid: what the inference renderer needs
    (app/nerf/test.py:53-94)."""
    import torch

    latents = model.position_encoder.latents
    table, flat = flat_from_nerf_param_tree(state["params"]["nerf"], *latents.shape)
    with torch.no_grad():
        latents.copy_(torch.from_numpy(table).to(latents.device))
        model.mlp_flat[: nerf_mod.MLP_NUMEL].copy_(torch.from_numpy(flat).to(model.mlp_flat.device))
    if "ogrid" not in state:
        return None
    return torch.from_numpy(np.ascontiguousarray(state["ogrid"]["occupancy"]).astype(np.uint8, copy=False)).to(latents.device)


# ---------------------------------------------------------------------------------------------- .npz container
def _flatten(tree, prefix=""):
    for k, v in tree.items():
        key = f"{prefix}{k}"
        if isinstance(v, dict):
            yield from _flatten(v, key + "/")
        else:
            yield key, v


def save_npz(path, state: dict):
    """One array per leaf, keys joined with '/', ``None`` leaves recorded in ``__none__``."""
    arrays, nones = {}, []
    for key, v in _flatten(state):
        if v is None:
            nones.append(key)
        else:
            arrays[key] = np.asarray(v)
    arrays["__none__"] = np.array(nones, dtype=np.str_)
    np.savez(path, **arrays)


def load_npz(path) -> dict:
    state = {}
    with np.load(path, allow_pickle=False) as z:
        entries = [(k, z[k]) for k in z.files if k != "__none__"] + [(str(k), None) for k in z["__none__"]]
    for key, v in entries:
        node = state
        *parents, leaf = key.split("/")
        for p in parents:
            node = node.setdefault(p, {})
        node[leaf] = v if v is None or v.ndim else v.item()
    return state


# ---------------------------------------------------------------------------------------------- flax msgpack container
_EXT_NDARRAY, _EXT_NPSCALAR = 1, 3


def _msgpack():
    try:
        import msgpack
    except ImportError as exc:  # pragma: no cover - the image ships it
        raise CheckpointError("the flax container needs the `msgpack` package") from exc
    return msgpack


def _pack_array(a):
    a = np.asarray(a)
    if a.dtype.hasobject:
        raise CheckpointError("object arrays cannot be serialised")
    return _msgpack().packb((a.shape, a.dtype.name, a.tobytes("C")), use_bin_type=True)


def _ext_pack(x):
    if isinstance(x, np.ndarray):
        return _msgpack().ExtType(_EXT_NDARRAY, _pack_array(x))
    if isinstance(x, np.generic):
        return _msgpack().ExtType(_EXT_NPSCALAR, _pack_array(x))
    raise CheckpointError(f"cannot serialise a {type(x).__name__} leaf")


def _ext_unpack(code, data):
    if code in (_EXT_NDARRAY, _EXT_NPSCALAR):
        shape, dtype_name, buf = _msgpack().unpackb(data, raw=True)
        a = np.frombuffer(buf, dtype=np.dtype(dtype_name.decode())).reshape(shape)
        return a if code == _EXT_NDARRAY else a[()]
    return _msgpack().ExtType(code, data)


def _indexed(tree):
    """flax stores tuples and lists as dicts keyed by their index."""
    if isinstance(tree, dict):
        return {str(k): _indexed(v) for k, v in tree.items()}
    if isinstance(tree, (list, tuple)):
        return {str(i): _indexed(v) for i, v in enumerate(tree)}
    return tree


def msgpack_serialize(state: dict) -> bytes:
    """``flax.serialization.msgpack_serialize`` for trees of dict / None / int / float / str / numpy leaves (arrays below
    flax's 1 GiB chunking limit, which covers every table this path supports)."""
    for key, v in _flatten(_indexed(state)):
        if isinstance(v, np.ndarray) and v.nbytes >= 2 ** 30:
            raise CheckpointError(f"{key}: arrays of 1 GiB and more need flax's chunked layout, not written here")
    return _msgpack().packb(_indexed(state), default=_ext_pack, strict_types=True)


def msgpack_restore(data: bytes) -> dict:
    return _msgpack().unpackb(data, ext_hook=_ext_unpack, raw=False, strict_map_key=False)


def save_flax_checkpoint(ckpt_dir, state: dict, prefix="checkpoint_") -> str:
    """Writes ``<ckpt_dir>/<prefix><step>`` the way ``flax.training.checkpoints.save_checkpoint`` names it.  The private
    ``opt_state_b200`` entry is left out: a flax ``restore_checkpoint(target=...)`` rejects fields its target lacks."""
    os.makedirs(ckpt_dir, exist_ok=True)
    path = os.path.join(ckpt_dir, f"{prefix}{int(state['step'])}")
    body = {k: v for k, v in state.items() if k != "opt_state_b200"}
    body.setdefault("opt_state", None)
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(msgpack_serialize(body))
    os.replace(tmp, path)
    return path


def load_flax_checkpoint(path) -> dict:
    """``path``: a checkpoint file, or a directory -- then the file with the highest step is taken, as
    ``restore_checkpoint`` does."""
    if os.path.isdir(path):
        steps = []
        for name in os.listdir(path):
            head, _, tail = name.rpartition("_")
            if head and tail.isdigit():
                steps.append((int(tail), name))
        if not steps:
            raise CheckpointError(f"no checkpoint_<step> file in {path}")
        path = os.path.join(path, max(steps)[1])
    with open(path, "rb") as f:
        return msgpack_restore(f.read())

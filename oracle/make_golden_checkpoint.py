"""TEST INFRASTRUCTURE ONLY -- the reference's checkpoint tree, read off its OWN code.

What a checkpoint of the reference holds is decided by four pieces of its source, none of which needs jax to be read:

  * ``app/nerf/train.py:172-199``   the ``params={...}`` dict handed to ``NeRFState.create``            (AST)
  * ``utils/types.py:93-144``       the pytree-node fields of ``OccupancyDensityGrid``                  (AST)
  * ``app/nerf/_utils.py:64-76``    the weight-decay mask, which spells the sub-module names of ``nerf`` (AST)
  * ``models/nerfs.py`` / ``models/encoders.py``  the parameter names and shapes the modules ask flax for (executed
    unmodified on numpy through oracle/ref_shim.py, with every ``self.param(...)`` / ``nn.Dense`` request logged)

This script exports a model with ``jaxngp_b200/checkpoint.py``, binds the exported tree -- by ITS names -- to the
reference's ``make_nerf_ngp`` model, checks that the model reproduces the committed outputs of
tests/golden/nerf_reference.npz bit for bit, and writes tests/golden/checkpoint_reference.json (names, shapes, dtypes).

    python oracle/make_golden_checkpoint.py        # needs /root/reference; run in the build container only
"""
import ast
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REFERENCE = "/root/reference"


def _dict_keys(node):
    return [k.value for k in node.keys if isinstance(k, ast.Constant)]


def params_keys():
    """Keys of the ``params=`` dict of ``NeRFState.create(...)`` in app/nerf/train.py."""
    tree = ast.parse(open(os.path.join(REFERENCE, "app", "nerf", "train.py")).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr == "create" \
                and getattr(node.func.value, "id", None) == "NeRFState":
            for kw in node.keywords:
                if kw.arg == "params":
                    return _dict_keys(kw.value)
    raise RuntimeError("NeRFState.create(params=...) not found")


def ogrid_fields():
    """Annotated fields of OccupancyDensityGrid that are pytree nodes (no ``struct.field(pytree_node=False)``)."""
    tree = ast.parse(open(os.path.join(REFERENCE, "utils", "types.py")).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "OccupancyDensityGrid")
    out = []
    for n in cls.body:
        if isinstance(n, ast.AnnAssign):
            static = n.value is not None and "pytree_node=False" in ast.unparse(n.value)
            if not static:
                out.append(n.target.id)
    return out


def weight_decay_mask():
    """The ``mask=`` dict of ``optax.add_decayed_weights`` in make_optimizer: {"nerf": {submodule: bool}, ...}."""
    tree = ast.parse(open(os.path.join(REFERENCE, "app", "nerf", "_utils.py")).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and getattr(node.func, "attr", None) == "add_decayed_weights":
            for kw in node.keywords:
                if kw.arg == "mask":
                    return ast.literal_eval(kw.value)
    raise RuntimeError("add_decayed_weights(mask=...) not found")


def main():
    from jaxngp_b200 import checkpoint as C
    from jaxngp_b200 import nerf as nerf_mod
    from oracle import hashgrid_np as H
    from oracle import ref_shim
    from tests import inputs

    nerfs = ref_shim.install_nerf()
    requested = {}  # "<submodule>/<leaf path>" -> (shape, dtype) as the reference's modules ask for them
    current = []

    orig_param = ref_shim.Module.param

    def logging_param(self, name, init_fn, shape, dtype=np.float32):
        requested[f"{current[-1]}/{name}"] = (list(map(int, shape)), np.dtype(dtype).name)
        return orig_param(self, name, init_fn, shape, dtype)

    orig_dense_call = ref_shim.Dense.__call__

    def logging_dense(self, x):
        parent = ref_shim._module_stack[-1]
        i = parent._auto_index.get("Dense", 0)
        requested[f"{current[-1]}/Dense_{i}/kernel"] = ([int(x.shape[-1]), int(self.features)], "float32")
        return orig_dense_call(self, x)

    ref_shim.Module.param = logging_param
    ref_shim.Dense.__call__ = logging_dense

    g = np.load(os.path.join(ROOT, "tests", "golden", "nerf_reference.npz"))
    lv = H.level_table(16, 2 ** 19, 2, 16, 2048, 3)
    rows = int(lv["offsets"][-1])
    table = inputs.encoder_table(rows, 2, amp=1.0)
    flat = np.concatenate([g[name].reshape(-1) for name, _, _ in nerf_mod.MLP_SHAPES]).astype(np.float32)

    class Grid:
        density = np.zeros(128 ** 3, np.float32)
        occ_mask = np.zeros(128 ** 3, bool)
        occupancy = np.full(128 ** 3 // 8, 255, np.uint8)

    state = C.make_state(7, table, flat, Grid, n_frames=100)
    tree = state["params"]["nerf"]

    # bind the exported tree to the reference's model by the exported names; sub-modules are the attributes of NeRF
    model = nerfs.make_nerf_ngp(bound=1.0, inference=False)
    for sub, leaves in tree.items():
        module = getattr(model, sub)  # AttributeError = a sub-module name the reference does not have
        module.bind_params(**{k: (v["kernel"] if isinstance(v, dict) else v) for k, v in leaves.items()})

    # log which sub-module is running: wrap the bound sub-modules' __call__ through the parent's attribute access
    for sub in tree:
        module = getattr(model, sub)
        cls = type(module)
        if not getattr(cls, "_ckpt_logged", False):
            inner = cls.__call__

            def make(inner):
                def call(self, *a, **k):
                    name = next(s for s in tree if getattr(model, s) is self)
                    current.append(name)
                    try:
                        return inner(self, *a, **k)
                    finally:
                        current.pop()
                return call

            cls.__call__ = make(inner)
            cls._ckpt_logged = True

    drgbs, _ = model(g["xyz"], g["dirs"], np.zeros((0,), np.float32))
    assert np.array_equal(np.asarray(drgbs), g["drgbs"]), "the exported tree does not reproduce the reference model's outputs"

    exported = {key[len("params/nerf/"):]: (list(v.shape), v.dtype.name)
                for key, v in C._flatten(state) if key.startswith("params/nerf/")}
    assert exported == requested, (exported, requested)

    mask = weight_decay_mask()
    assert set(mask["nerf"]) == set(tree), (mask, list(tree))
    out = {"params_keys": params_keys(), "ogrid_fields": ogrid_fields(), "weight_decay_mask": mask,
           "nerf_leaves_requested_by_the_reference": {k: {"shape": s, "dtype": d} for k, (s, d) in sorted(requested.items())},
           "appearance_embeddings_shape": [100, 0]}
    assert set(out["params_keys"]) == set(state["params"]), (out["params_keys"], list(state["params"]))
    assert set(out["ogrid_fields"]) == set(state["ogrid"]), (out["ogrid_fields"], list(state["ogrid"]))
    path = os.path.join(ROOT, "tests", "golden", "checkpoint_reference.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", path)
    print(json.dumps(out, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()

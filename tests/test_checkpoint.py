"""Checkpoint / array interchange with the reference (jaxngp_b200/checkpoint.py), CPU only.

tests/golden/checkpoint_reference.json was read off the reference's OWN code by oracle/make_golden_checkpoint.py: the
``params`` keys of ``NeRFState.create`` (app/nerf/train.py:187-197), the pytree-node fields of ``OccupancyDensityGrid``
(utils/types.py:93-107), the sub-module names of the weight-decay mask (app/nerf/_utils.py:64-76) and every parameter
name / shape that ``make_nerf_ngp``'s modules request while running unmodified on numpy -- the same run reproduced the
model outputs of tests/golden/nerf_reference.npz bit for bit from the exported tree."""
import json
import os

import numpy as np
import pytest

from jaxngp_b200 import checkpoint as C
from jaxngp_b200 import nerf as nerf_mod

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROWS, F, G3 = 6098120, 2, 128 ** 3


class _Grid:
    def __init__(self, rng, n=G3):
        self.density = rng.uniform(0, 3, n).astype(np.float32)
        self.occ_mask = self.density > 1.5
        self.occupancy = np.packbits(self.occ_mask, bitorder="little")


def _state(rng, rows=ROWS, with_moments=False, n=G3):
    table = rng.uniform(-1e-4, 1e-4, (rows, F)).astype(np.float32)
    flat = rng.uniform(-0.3, 0.3, nerf_mod.MLP_NUMEL).astype(np.float32)
    kw = {}
    if with_moments:
        kw = dict(adam_m=rng.normal(size=rows * F + nerf_mod.MLP_NUMEL + 24).astype(np.float32),
                  adam_v=rng.uniform(size=rows * F + nerf_mod.MLP_NUMEL + 24).astype(np.float32))
    return C.make_state(1234, table, flat, _Grid(rng, n), n_frames=100, **kw), table, flat


def _leaves(tree):
    return {k: v for k, v in C._flatten(tree)}


def test_checkpoint_tree_binds_to_the_reference_model():
    ref = json.load(open(os.path.join(GOLDEN, "checkpoint_reference.json")))
    state, table, flat = _state(np.random.default_rng(0))
    assert set(state["params"]) == set(ref["params_keys"])
    assert set(state["ogrid"]) == set(ref["ogrid_fields"])
    assert set(state["params"]["nerf"]) == set(ref["weight_decay_mask"]["nerf"])
    got = {k: {"shape": list(v.shape), "dtype": v.dtype.name} for k, v in _leaves(state["params"]["nerf"]).items()}
    assert got == ref["nerf_leaves_requested_by_the_reference"]
    assert state["params"]["bg"] is None  # train.py:189 when scene_meta.bg is false
    assert list(state["params"]["appearance_embeddings"].shape) == ref["appearance_embeddings_shape"]
    assert state["ogrid"]["alive_indices"].dtype == np.uint32 and state["ogrid"]["occ_mask"].dtype == np.bool_
    # weights in the [in, out] orientation of flax's Dense, in the flat buffer's order
    off = 0
    for (module, layer, name), (nm, i, o) in zip(C._MLP_TREE, nerf_mod.MLP_SHAPES):
        assert name == nm
        assert np.array_equal(state["params"]["nerf"][module][layer]["kernel"], flat[off:off + i * o].reshape(i, o))
        off += i * o
    t2, f2 = C.flat_from_nerf_param_tree(state["params"]["nerf"], ROWS, F)
    assert np.array_equal(t2, table) and np.array_equal(f2, flat)


@pytest.mark.parametrize("container", ["npz", "flax"])
def test_checkpoint_round_trip(tmp_path, container):
    rng = np.random.default_rng(1)
    state, _, _ = _state(rng, rows=4096, with_moments=(container == "npz"), n=32 ** 3)
    if container == "npz":
        path = str(tmp_path / "state.npz")
        C.save_npz(path, state)
        back = C.load_npz(path)
    else:
        path = C.save_flax_checkpoint(str(tmp_path), state)
        assert os.path.basename(path) == "checkpoint_1234"  # flax.training.checkpoints naming
        C.save_flax_checkpoint(str(tmp_path), dict(state, step=99))
        back = C.load_flax_checkpoint(str(tmp_path))  # a directory: the highest step wins
        assert "opt_state_b200" not in back
    assert int(back["step"]) == 1234 and back["params"]["bg"] is None
    a, b = _leaves(state), _leaves(back)
    if container == "flax":
        a = {k: v for k, v in a.items() if not k.startswith("opt_state_b200")}
    assert set(a) == set(b)
    for k, v in a.items():
        if isinstance(v, np.ndarray):
            assert b[k].dtype == v.dtype and b[k].shape == v.shape and np.array_equal(b[k], v), k


def test_flax_wire_format_known_answers():
    """The container against flax.serialization's published encoding, byte for byte on a small tree."""
    import msgpack
    blob = C.msgpack_serialize({"step": 3, "t": (np.arange(3, dtype=np.uint8), None), "s": np.float32(1.5)})
    top = msgpack.unpackb(blob, raw=False, strict_map_key=False)
    assert top["step"] == 3 and set(top["t"]) == {"0", "1"} and top["t"]["1"] is None  # tuples -> index-keyed dicts
    arr, scalar = top["t"]["0"], top["s"]
    assert (arr.code, scalar.code) == (1, 3)  # _MsgpackExtType.ndarray / npscalar
    assert msgpack.unpackb(arr.data, raw=True) == [[3], b"uint8", b"\x00\x01\x02"]
    assert msgpack.unpackb(scalar.data, raw=True) == [[], b"float32", np.float32(1.5).tobytes()]
    back = C.msgpack_restore(blob)
    assert back["s"] == np.float32(1.5) and isinstance(back["s"], np.float32) and np.array_equal(back["t"]["0"], [0, 1, 2])


def test_checkpoint_rejects_what_the_path_does_not_support():
    state, _, _ = _state(np.random.default_rng(2), rows=512, n=16 ** 3)
    tree = state["params"]["nerf"]
    with pytest.raises(C.CheckpointError):  # another table geometry
        C.flat_from_nerf_param_tree(tree, rows=1024, F=2)
    wide = {**tree, "rgb_mlp": {**tree["rgb_mlp"], "Dense_0": {"kernel": np.zeros((40, 64), np.float32)}}}
    with pytest.raises(C.CheckpointError):  # appearance embeddings widen the colour MLP's input (nerfs.py:79-83)
        C.flat_from_nerf_param_tree(wide)
    deep = {**tree, "density_mlp": {**tree["density_mlp"], "Dense_2": {"kernel": np.zeros((16, 16), np.float32)}}}
    with pytest.raises(C.CheckpointError):
        C.flat_from_nerf_param_tree(deep)
    with pytest.raises(C.CheckpointError):  # a frequency-encoded model has no hash table
        C.flat_from_nerf_param_tree({k: v for k, v in tree.items() if k != "position_encoder"})
    half = {**tree, "position_encoder": {C.TABLE_NAME: tree["position_encoder"][C.TABLE_NAME].astype(np.float16)}}
    with pytest.raises(C.CheckpointError):
        C.flat_from_nerf_param_tree(half)
    with pytest.raises(C.CheckpointError):
        C.load_flax_checkpoint(os.path.dirname(os.path.abspath(__file__)))  # no checkpoint_<step> file there


def test_checkpoint_loads_into_the_host_model():
    """Parameters travel reference tree -> this package's model buffers in place (no kernel runs: host I/O only)."""
    import torch
    model = nerf_mod.NeRF(bound=1.0, inference=True, device="cpu", T=1 << 12)
    rows, f = model.position_encoder.latents.shape
    rng = np.random.default_rng(3)
    state, table, flat = _state(rng, rows=rows, n=16 ** 3)
    ptr = model.position_encoder.latents.data_ptr()
    bitfield = C.load_into_model(model, state)
    assert model.position_encoder.latents.data_ptr() == ptr  # in place: captured graphs keep their pointers
    assert np.array_equal(model.position_encoder.latents.detach().numpy(), table)
    assert np.array_equal(model.mlp_flat.detach().numpy()[: nerf_mod.MLP_NUMEL], flat)
    assert np.array_equal(model.rgb_w2.detach().numpy(), state["params"]["nerf"]["rgb_mlp"]["Dense_2"]["kernel"])
    assert bitfield.dtype == torch.uint8 and np.array_equal(bitfield.numpy(), state["ogrid"]["occupancy"])
    back = C.make_state(5, model.position_encoder.latents, model.mlp_flat)
    assert np.array_equal(back["params"]["nerf"]["density_mlp"]["Dense_1"]["kernel"],
                          state["params"]["nerf"]["density_mlp"]["Dense_1"]["kernel"]) and "ogrid" not in back
    with pytest.raises(C.CheckpointError):
        C.load_into_model(nerf_mod.NeRF(bound=1.0, inference=True, device="cpu", T=1 << 10), state)


def test_adam_moments_are_found_in_a_reference_written_opt_state():
    """optax's wrappers differ between versions; the network optimizer's ScaleByAdamState is found wherever they put it
    (here: chain -> multi_transform inner_states -> masked inner_state -> chain, tuples as index-keyed dicts)."""
    state, _, _ = _state(np.random.default_rng(4), rows=256, n=16 ** 3)
    assert "opt_state" in state and state["opt_state"] is None  # the field flax's restore insists on
    mu = {"nerf": C.nerf_param_tree(np.ones((256, 2), np.float32), np.full(nerf_mod.MLP_NUMEL, 2, np.float32)), "bg": None,
          "appearance_embeddings": {}}
    nu = {"nerf": C.nerf_param_tree(np.full((256, 2), 3, np.float32), np.full(nerf_mod.MLP_NUMEL, 4, np.float32)), "bg": None,
          "appearance_embeddings": {}}
    ae = {"count": np.int32(9), "mu": {"nerf": {}, "bg": None, "appearance_embeddings": np.zeros((100, 0), np.float32)},
          "nu": {"nerf": {}, "bg": None, "appearance_embeddings": np.zeros((100, 0), np.float32)}}
    opt_state = {"0": {"inner_states": {"ae": {"inner_state": {"0": ae, "1": {}}},
                                        "network": {"inner_state": {"0": {"count": np.int32(9), "mu": mu, "nu": nu},
                                                                    "1": {"count": np.int32(9)}}}}},
                 "1": {"inner_state": {}}}
    m, v, count = C.find_adam_moments(opt_state)
    assert m is mu["nerf"] and v is nu["nerf"] and int(count) == 9
    assert C.find_adam_moments({"0": {"inner_state": {}}}) is None and C.find_adam_moments(None) is None
    back = C.msgpack_restore(C.msgpack_serialize(dict(state, opt_state=opt_state)))
    m2, v2, _ = C.find_adam_moments(back["opt_state"])
    assert np.array_equal(m2["rgb_mlp"]["Dense_2"]["kernel"], mu["nerf"]["rgb_mlp"]["Dense_2"]["kernel"])
    assert np.array_equal(v2["position_encoder"][C.TABLE_NAME], nu["nerf"]["position_encoder"][C.TABLE_NAME])


def test_checkpoint_round_trip_through_trainer_buffers():
    """state_from_trainer / load_into_trainer on a stand-in that has the Trainer's buffer attributes (flat [table | MLP]
    moments, device step counter, density grid) on the CPU: no kernel is involved in a load or a snapshot."""
    import types
    import torch
    rows, n_cells = 512, 16 ** 3
    rng = np.random.default_rng(5)
    n_table = rows * F
    total = -(-(n_table + nerf_mod.MLP_NUMEL) // 32) * 32

    def trainer(seed):
        g = torch.Generator().manual_seed(seed)
        flat_params = torch.rand(total, generator=g)
        grid = types.SimpleNamespace(density=torch.rand(n_cells, generator=g), occ_mask=torch.rand(n_cells, generator=g) > 0.5,
                                     occupancy=torch.randint(0, 256, (n_cells // 8,), generator=g, dtype=torch.uint8))
        return types.SimpleNamespace(
            world_size=1, device=torch.device("cpu"), table=flat_params[:n_table].view(rows, F),
            mlp_flat=flat_params[n_table:n_table + nerf_mod.MLP_NUMEL], table_numel=n_table,
            n_params=n_table + nerf_mod.MLP_NUMEL, adam_m=torch.rand(total, generator=g), adam_v=torch.rand(total, generator=g),
            grid=grid, step=seed, step_dev=torch.full((1,), seed, dtype=torch.int32), scene=types.SimpleNamespace(n_views=100),
            _prefetched=("stale", 0), flat_params=flat_params)

    src, dst = trainer(41), trainer(7)
    state = C.state_from_trainer(src)
    assert state["step"] == 41 and state["params"]["appearance_embeddings"].shape == (100, 0)
    C.load_into_trainer(dst, state)
    assert torch.equal(dst.flat_params[: dst.n_params], src.flat_params[: src.n_params])
    assert torch.equal(dst.adam_m[: dst.n_params], src.adam_m[: src.n_params]) and torch.equal(dst.adam_v[: dst.n_params], src.adam_v[: src.n_params])
    for name in ("density", "occ_mask", "occupancy"):
        assert torch.equal(getattr(dst.grid, name), getattr(src.grid, name))
    assert dst.step == 41 and int(dst.step_dev) == 41 and dst._prefetched is None
    # sharded optimizer state is refused rather than half-loaded
    dst.world_size = 2
    before = dst.flat_params.clone()
    with pytest.raises(C.CheckpointError):
        C.load_into_trainer(dst, dict(state, params=dict(state["params"], nerf=C.nerf_param_tree(np.zeros((rows, F), np.float32), np.zeros(nerf_mod.MLP_NUMEL, np.float32)))))
    assert torch.equal(dst.flat_params, before)  # validated before the first copy
    with pytest.raises(C.CheckpointError):
        C.state_from_trainer(dst)
    dst.world_size = 1
    with pytest.raises(C.CheckpointError):  # a grid of another resolution
        C.load_into_trainer(dst, dict(state, ogrid={k: np.concatenate([v, v]) for k, v in state["ogrid"].items()}))
    with pytest.raises(C.CheckpointError):  # background model
        C.load_into_trainer(dst, dict(state, params=dict(state["params"], bg={"Dense_0": {}})))


@pytest.mark.gpu
def test_trainer_resumes_bit_identically_from_a_flax_checkpoint(tmp_path):
    """Train, write a flax-format checkpoint (+ the .npz with the moments), load them into fresh trainers: parameters,
    moments and grid arrive as the same bits, and the next step sees the same forward."""
    import torch
    from jaxngp_b200.trainer import Scene, Trainer
    dev = "cuda:0"
    scene = Scene(dev, n_views=4, width=100, height=100)
    n_rays = 1 << 12

    def perm(seed):
        g = torch.Generator(device=dev).manual_seed(seed)
        return torch.randint(0, scene.n_pixels, (n_rays,), device=dev, generator=g, dtype=torch.int32)

    a = Trainer(device=dev, n_rays=n_rays, total_samples=1 << 15, scene=scene, use_graph=False)
    for it in range(20):
        a.train_step(perm(it))
    a.update_ogrid()
    state = C.state_from_trainer(a)
    path = C.save_flax_checkpoint(str(tmp_path), state)
    C.save_npz(str(tmp_path / "state.npz"), state)
    from_flax, from_npz = C.load_flax_checkpoint(path), C.load_npz(str(tmp_path / "state.npz"))
    assert int(from_flax["step"]) == 20 and "opt_state_b200" in from_npz

    b = Trainer(device=dev, n_rays=n_rays, total_samples=1 << 15, scene=scene, use_graph=False, seed=5)
    C.load_into_trainer(b, from_npz)
    assert torch.equal(b.flat_params[: b.n_params], a.flat_params[: a.n_params])
    assert torch.equal(b.adam_m[: b.n_params], a.adam_m[: a.n_params]) and torch.equal(b.grid.occupancy, a.grid.occupancy)
    noises = torch.rand(n_rays, device=dev)
    bg = torch.rand(n_rays, 3, device=dev)
    # same parameters, grid and inputs -> same forward (the backward's atomic scatter order is free, so the parameters
    # after the step are not compared bit for bit)
    outs = [t._step_body(perm(99), noises, bg) for t in (a, b)]
    assert abs(float(outs[0]["loss"]) - float(outs[1]["loss"])) <= 1e-6 * abs(float(outs[0]["loss"]))
    assert int(outs[0]["measured_batch_size"]) == int(outs[1]["measured_batch_size"])
    # the flax file alone (no moments) restores the model: same rendering inputs
    c = Trainer(device=dev, n_rays=n_rays, total_samples=1 << 15, scene=scene, use_graph=False, seed=6)
    before = C.state_from_trainer(b)["params"]["nerf"]
    C.load_into_trainer(c, from_flax)
    assert int(c.step_dev.item()) == 20 and float(c.adam_m.abs().max()) == 0  # moments stay at their initial zeros
    after = C.state_from_trainer(c)["params"]["nerf"]
    assert np.array_equal(after["rgb_mlp"]["Dense_1"]["kernel"], state["params"]["nerf"]["rgb_mlp"]["Dense_1"]["kernel"])
    assert not np.array_equal(after["rgb_mlp"]["Dense_1"]["kernel"], before["rgb_mlp"]["Dense_1"]["kernel"])  # b has stepped since
    # culled cells survive the round trip: the trainable-cell table comes back, so a full update leaves them at -1
    culled = torch.zeros_like(a.grid.density, dtype=torch.bool)
    culled[::3] = True
    a.grid.density[culled] = -1.0
    a.grid.alive_indices = torch.nonzero(~culled).reshape(-1).to(torch.int32)
    a.grid.alive_indices_offset = [0, int((~culled).sum())]
    d = Trainer(device=dev, n_rays=n_rays, total_samples=1 << 15, scene=scene, use_graph=False, seed=7)
    C.load_into_trainer(d, C.state_from_trainer(a))
    assert torch.equal(d.grid.alive_indices, a.grid.alive_indices) and d.grid.alive_indices_offset == a.grid.alive_indices_offset
    d.update_ogrid(update_all=True)
    assert torch.equal(d.grid.density < 0, culled)

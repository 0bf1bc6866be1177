"""The drop-in boundary against the reference's OWN binding code, CPU only.

tests/golden/lowering_reference.json was recorded by oracle/make_golden_lowering.py, which ran -- unmodified -- all ten
``*_lowering_rule`` functions of volume-rendering-jax and jax-tcnn (against recording stand-ins for the few MLIR
constructors they use, with this repo's drop-in extension modules where the reference's compiled ones go) and the
registration loops of their ``impl.py`` files, and read the module surface off the two ``ffi.cc`` files.  Here:

  * the drop-in modules export what ffi.cc exports, and every capsule carries the address of the matching C symbol;
  * for the same operand shapes, the host mirror hands libngp_b200.so exactly the buffers the reference's lowering
    hands XLA -- operands in order, then results in order, same shapes, same element widths -- and the same opaque
    bytes, under the symbol ``ngp_<target>``.  Nothing is launched: ``_lib.call`` is replaced by a recorder.
"""
import ctypes
import json
import os

import pytest
import torch

from jaxngp_b200 import _lib
from jaxngp_b200 import jaxtcnn
from jaxngp_b200 import volrendjax as V
from jaxngp_b200.jaxtcnn import tcnnutils
from jaxngp_b200.volrendjax import integrating, volrendutils_cuda

GOLDEN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lowering_reference.json")))
CALLS = GOLDEN["custom_calls"]

# element types as the reference's lowering names them -> what the torch host carries (uint32 travels as int32 bits)
TORCH = {"float32": torch.float32, "uint32": torch.int32, "uint8": torch.uint8, "bool": torch.bool}


def _capsule_pointer(capsule):
    get_ptr = ctypes.pythonapi.PyCapsule_GetPointer
    get_ptr.restype, get_ptr.argtypes = ctypes.c_void_p, [ctypes.py_object, ctypes.c_char_p]
    get_name = ctypes.pythonapi.PyCapsule_GetName
    get_name.restype, get_name.argtypes = ctypes.c_char_p, [ctypes.py_object]
    name = get_name(capsule)
    return name, get_ptr(capsule, name)


@pytest.mark.parametrize("package,module", [("volrendjax", volrendutils_cuda), ("jaxtcnn", tcnnutils)])
def test_drop_in_extension_modules_export_what_ffi_cc_exports(built_lib, package, module):
    ffi = GOLDEN["ffi"][package]
    assert module.__name__.rsplit(".", 1)[1] == ffi["module"]  # importable as `from .. import <module>` in the reference's tree
    for fn in ffi["functions"]:
        assert callable(getattr(module, fn)), fn
    L = built_lib
    for getter, names in ffi["registrations"].items():
        regs = getattr(module, getter)()
        assert list(regs) == names
        for name, capsule in regs.items():
            cap_name, ptr = _capsule_pointer(capsule)
            assert cap_name.decode() == ffi["capsule_name"] == "xla._CUSTOM_CALL_TARGET"
            assert ptr == ctypes.cast(getattr(L, "ngp_" + name), ctypes.c_void_p).value and ptr
    registered = {n for n, platform in GOLDEN["registered_by_the_reference_impl"] if platform == "gpu"}
    assert registered == set(CALLS) and len(GOLDEN["registered_by_the_reference_impl"]) == len(CALLS)


def test_descriptor_factories_reject_what_ffi_cc_rejects():
    with pytest.raises(RuntimeError, match="expected n_bytes to be a positive integer, got 0"):  # ffi.cc:57-59
        volrendutils_cuda.make_packbits_descriptor(0)
    with pytest.raises(RuntimeError, match="expected K to be a positive integer, got 0"):  # ffi.cc:79-81
        volrendutils_cuda.make_marching_descriptor(1, 1, 1, 0, 1, 1.0, 0.0)
    with pytest.raises(RuntimeError, match="expected K to be a positive integer, got 0"):  # ffi.cc:114-116
        volrendutils_cuda.make_marching_inference_descriptor(1, 1, 1, 0, 1, 1, 1.0, 0.0)


def _operands(target):
    """CPU tensors shaped like the recorded operands, by the rule's own parameter names."""
    c = CALLS[target]
    out = {}
    for name, (shape, dtype) in c["rule_arguments"].items():  # every array argument of the rule (it may not pass all on)
        out[name] = torch.zeros(shape, dtype=TORCH[dtype])
    return out, c["statics"]


def _drive(target):
    a, st = _operands(target)
    if target == "pack_density_into_bits":
        V.packbits(a["density_threshold"] + 0.5, a["density_grid"])
    elif target == "morton3d":
        V.morton3d(a["xyzs"])
    elif target == "morton3d_invert":
        V.morton3d_invert(a["idcs"])
    elif target == "march_rays":
        V.march_rays(**st, **a)
    elif target == "march_rays_inference":
        a["indices"] = a.pop("indices_in")  # the public wrapper's name (marching/__init__.py:96-110)
        V.march_rays_inference(**st, **a)
    elif target == "integrate_rays":
        integrating._integrate_fwd(a["rays_sample_startidx"], a["rays_n_samples"], a["bgs"], a["dss"], a["z_vals"], a["drgbs"])
    elif target == "integrate_rays_backward":
        integrating._integrate_bwd(st["near_distance"], a["rays_sample_startidx"], a["rays_n_samples"], a["bgs"], a["dss"],
                                   a["z_vals"], a["drgbs"], a["final_rgbds"], a["final_opacities"], a["dL_dfinal_rgbds"])
    elif target == "integrate_rays_inference":
        V.integrate_rays_inference(**a)
    elif target in ("hashgrid_encode", "hashgrid_encode_backward"):
        desc = jaxtcnn.HashGridMetadata(L=st["L"], F=st["F"], N_min=st["N_min"], per_level_scale=st["per_level_scale"])
        params = a["params"].requires_grad_(True)
        enc = jaxtcnn.hashgrid_encode(desc, a["offset_table_data"], a["coords_rm"], params)
        if target == "hashgrid_encode_backward":
            enc.backward(torch.ones_like(enc))
    else:
        raise KeyError(target)


@pytest.mark.parametrize("target", sorted(CALLS))
def test_host_mirror_issues_the_custom_call_the_reference_lowers_to(monkeypatch, target):
    seen = []

    def recorder(name, buffers, opaque, stream=None):
        seen.append((name, [(tuple(b.shape), b.dtype, b.is_contiguous()) for b in buffers], bytes(opaque)))

    monkeypatch.setattr(_lib, "call", recorder)
    _drive(target)
    mine = [s for s in seen if s[0] == "ngp_" + target]
    assert len(mine) == 1, [s[0] for s in seen]
    _, buffers, opaque = mine[0]
    ref = CALLS[target]
    expected = [(tuple(shape), TORCH[dtype]) for shape, dtype in ref["operands"] + ref["results"]]
    assert [(s, d) for s, d, _ in buffers] == expected
    assert all(contig for _, _, contig in buffers)
    # row-major everywhere: the reference asks XLA for its default (descending) layouts
    for layouts, tensors in ((ref["operand_layouts"], ref["operands"]), (ref["result_layouts"], ref["results"])):
        assert [list(x) for x in layouts] == [list(range(len(shape) - 1, -1, -1)) for shape, _ in tensors]
    assert opaque.hex() == ref["opaque"]


def test_integrate_rays_vjp_against_the_reference_rule(monkeypatch):
    """The reference's OWN bwd rule of integrate_rays (integrating/impl.py:111-141, run unmodified with named stand-ins):
    same residuals and cotangent handed to the backward primitive, in the same order -- and the documented difference
    (SURVEY quirk Q5): the reference returns ``dL_dbgs`` in the slot of ``dss`` and nothing for ``bgs``; here each
    cotangent goes to the argument it belongs to."""
    vjp = GOLDEN["integrate_rays_vjp"]
    assert vjp["backward_operands"] == ["rays_sample_startidx", "rays_n_samples", "bgs", "dss", "z_vals", "drgbs",
                                        "final_rgbds", "final_opacities", "dL_dfinal_rgbds"]
    assert vjp["backward_statics"] == ["near_distance"]
    assert vjp["cotangent_bound_to"] == {"near_distance": None, "rays_sample_startidx": None, "rays_n_samples": None,
                                         "bgs": None, "dss": "dL_dbgs", "z_vals": "dL_dz_vals", "drgbs": "dL_ddrgbs"}
    calls = []

    def recorder(name, buffers, opaque, stream=None):
        calls.append((name, [b.data_ptr() for b in buffers]))
        for b in buffers[-3:]:
            b.fill_(1.0)  # stand-in results so that autograd has something to route

    monkeypatch.setattr(_lib, "call", recorder)
    n, s = 6, 40
    start = torch.zeros(n, dtype=torch.int32)
    ns = torch.zeros(n, dtype=torch.int32)
    bgs = torch.zeros(n, 3, requires_grad=True)
    dss, z_vals = torch.zeros(s, requires_grad=True), torch.zeros(s, requires_grad=True)
    drgbs = torch.zeros(s, 4, requires_grad=True)
    _, final_rgbds, _ = V.integrate_rays(0.3, start, ns, bgs, dss, z_vals, drgbs)
    final_rgbds.sum().backward()
    fwd, bwd = calls
    assert (fwd[0], bwd[0]) == ("ngp_integrate_rays", "ngp_integrate_rays_backward")
    # backward operands = the forward's six inputs, its two ray outputs, then the cotangent: the reference's order
    assert bwd[1][:6] == fwd[1][:6] and bwd[1][6:8] == fwd[1][7:9] and len(bwd[1]) == 12
    assert bgs.grad is not None and bgs.grad.shape == (n, 3)      # dL_dbgs -> bgs (the reference: -> dss)
    assert dss.grad is None                                       # the op has no gradient for dss
    assert z_vals.grad.shape == (s,) and drgbs.grad.shape == (s, 4)


def test_status_returning_forms_and_jax_registration_shim(built_lib, monkeypatch):
    """The five-argument entry points (XLA's API_VERSION_STATUS_RETURNING) and ``jaxngp_b200/jax_ffi.py``: every drop-in
    target has a ``_status`` form behind a capsule of the right name; a failing call (descriptor of the wrong size: caught
    before anything touches a device) reports its message through the status callback exactly once, a null status pointer
    is tolerated, and ``register()`` hands all ten capsules to ``jax.ffi.register_ffi_target(..., api_version=0)`` -- here
    a recording stand-in, since jax is not installed."""
    import sys
    import types
    from jaxngp_b200 import jax_ffi
    L = built_lib
    caps = jax_ffi.capsules()
    assert sorted(caps) == sorted(CALLS)  # the reference's ten registered targets
    for target, capsule in caps.items():
        name, ptr = _capsule_pointer(capsule)
        assert name == b"xla._CUSTOM_CALL_TARGET"
        assert ptr == ctypes.cast(getattr(L, f"ngp_{target}_status"), ctypes.c_void_p).value and ptr
    seen = []
    CB = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t)
    cb = CB(lambda status, msg, n: seen.append((status, msg[:n].decode())))
    L.ngp_b200_set_status_failure_fn(ctypes.cast(cb, ctypes.c_void_p))
    try:
        bufs = (ctypes.c_void_p * 16)()
        token = ctypes.c_uint64(0)  # stands in for the XlaCustomCallStatus object
        L.ngp_march_rays_status(None, bufs, b"\x00" * 5, 5, ctypes.byref(token))
        assert len(seen) == 1 and seen[0][0] == ctypes.addressof(token)
        assert "march_rays: invalid opaque object size, expected 28, got 5" in seen[0][1]  # serde.h:35-40's message
        L.ngp_b200_clear_error()
        L.ngp_march_rays_status(None, bufs, b"\x00" * 5, 5, None)  # no status object: recorded, not reported
        assert len(seen) == 1 and L.ngp_b200_last_status() == -1
        L.ngp_b200_clear_error()
        # a well-formed call that has nothing to do reports nothing
        from jaxngp_b200 import descriptors as D
        desc = D.make_morton3d_descriptor(0)
        L.ngp_morton3d_status(None, bufs, desc, len(desc), ctypes.byref(token))
        assert len(seen) == 1 and L.ngp_b200_last_status() == 0
    finally:
        L.ngp_b200_set_status_failure_fn(None)
        L.ngp_b200_clear_error()
    # register() against a recording jax
    calls = []
    fake = types.ModuleType("jax")
    fake.ffi = types.SimpleNamespace(register_ffi_target=lambda name, capsule, platform, api_version: calls.append((name, platform, api_version, capsule)))
    monkeypatch.setitem(sys.modules, "jax", fake)
    assert jax_ffi.register() == sorted(CALLS)
    assert sorted(c[0] for c in calls) == sorted(CALLS) and {c[1:3] for c in calls} == {("CUDA", 0)}

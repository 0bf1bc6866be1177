"""TEST INFRASTRUCTURE ONLY -- the custom calls the reference emits, recorded from its OWN lowering rules.

Every native op of the reference reaches its kernel through a ``*_lowering_rule`` that builds the opaque descriptor and
emits ``custom_call(target, out_types, operands, backend_config=opaque, ...)``
(deps/volume-rendering-jax/src/volrendjax/{packbits,morton3d,marching,integrating}/lowering.py,
deps/jax-tcnn/src/jaxtcnn/hashgrid_tcnn/lowering.py).  Those files only touch a handful of MLIR type constructors, so
this script executes all ten rules UNMODIFIED against recording stand-ins for ``jax.interpreters.mlir`` /
``jaxlib.hlo_helpers`` -- with this repo's drop-in extension modules (jaxngp_b200/volrendjax/volrendutils_cuda.py,
jaxngp_b200/jaxtcnn/tcnnutils.py) in the place of the reference's compiled ones, exactly as a maintainer would install
them -- and writes what each rule handed to ``custom_call``: target name, operands and results (shape, dtype, in
order), layouts, and the opaque bytes.  It also runs the registration loops of the reference's ``impl.py`` files and
records the names they register, and reads the module surface off the two ``ffi.cc`` files.

    python oracle/make_golden_lowering.py        # needs /root/reference; writes tests/golden/lowering_reference.json

tests/test_bindings.py (CPU) then holds the host mirror to it: the buffers it passes to libngp_b200.so for the same
shapes are these operands followed by these results, with these opaque bytes, under the symbol ngp_<target>.
"""
import ast
import importlib.util
import json
import os
import re
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REFERENCE = "/root/reference"
VR = os.path.join(REFERENCE, "deps", "volume-rendering-jax")
TC = os.path.join(REFERENCE, "deps", "jax-tcnn")

# the operand signature of every rule at one small configuration (names = the rule's own parameter names)
N, S, K, G, CAP, NT, LEN, ROWS, L, F = 24, 160, 2, 16, 4, 40, 56, 1000, 16, 2
CASES = {
    "packbits_lowering_rule": dict(
        arrays={"density_threshold": ((K * G ** 3,), "float32"), "density_grid": ((K * G ** 3,), "float32")}, statics={}),
    "morton3d_lowering_rule": dict(arrays={"xyzs": ((LEN, 3), "uint32")}, statics={}),
    "morton3d_invert_lowering_rule": dict(arrays={"idcs": ((LEN,), "uint32")}, statics={}),
    "march_rays_lowering_rule": dict(
        arrays={"rays_o": ((N, 3), "float32"), "rays_d": ((N, 3), "float32"), "t_starts": ((N,), "float32"),
                "t_ends": ((N,), "float32"), "noises": ((N,), "float32"), "occupancy_bitfield": ((K * G ** 3 // 8,), "uint8")},
        statics=dict(total_samples=S, diagonal_n_steps=1024, K=K, G=G, bound=2.0, stepsize_portion=1 / 256)),
    "march_rays_inference_lowering_rule": dict(
        arrays={"rays_o": ((NT, 3), "float32"), "rays_d": ((NT, 3), "float32"), "t_starts": ((NT,), "float32"),
                "t_ends": ((NT,), "float32"), "occupancy_bitfield": ((K * G ** 3 // 8,), "uint8"),
                "next_ray_index_in": ((1,), "uint32"), "terminated": ((N,), "bool"), "indices_in": ((N,), "uint32")},
        statics=dict(diagonal_n_steps=1024, K=K, G=G, march_steps_cap=CAP, bound=2.0, stepsize_portion=1 / 256)),
    "integrate_rays_lowering_rule": dict(
        arrays={"rays_sample_startidx": ((N,), "uint32"), "rays_n_samples": ((N,), "uint32"), "bgs": ((N, 3), "float32"),
                "dss": ((S,), "float32"), "z_vals": ((S,), "float32"), "drgbs": ((S, 4), "float32")}, statics={}),
    "integrate_rays_backward_lowring_rule": dict(
        arrays={"rays_sample_startidx": ((N,), "uint32"), "rays_n_samples": ((N,), "uint32"), "bgs": ((N, 3), "float32"),
                "dss": ((S,), "float32"), "z_vals": ((S,), "float32"), "drgbs": ((S, 4), "float32"),
                "final_rgbds": ((N, 4), "float32"), "final_opacities": ((N,), "float32"), "dL_dfinal_rgbds": ((N, 4), "float32")},
        statics=dict(near_distance=0.3)),
    "integrate_rays_inference_lowering_rule": dict(
        arrays={"rays_bg": ((NT, 3), "float32"), "rays_rgbd": ((NT, 4), "float32"), "rays_T": ((NT,), "float32"),
                "n_samples": ((N,), "uint32"), "indices": ((N,), "uint32"), "dss": ((N, CAP), "float32"),
                "z_vals": ((N, CAP), "float32"), "drgbs": ((N, CAP, 4), "float32")}, statics={}),
    "hashgrid_encode_lowering_rule": dict(
        arrays={"offset_table_data": ((L + 1,), "uint32"), "coords_rm": ((3, LEN), "float32"), "params": ((ROWS, F), "float32")},
        statics=dict(L=L, F=F, N_min=16, per_level_scale=1.3819128799677762)),
    "hashgrid_encode_backward_lowering_rule": dict(
        arrays={"offset_table_data": ((L + 1,), "uint32"), "coords_rm": ((3, LEN), "float32"), "params": ((ROWS, F), "float32"),
                "dL_dy_rm": ((L * F, LEN), "float32"), "dy_dcoords_rm": ((3 * L * F, LEN), "float32")},
        statics=dict(L=L, F=F, N_min=16, per_level_scale=1.3819128799677762)),
}


class _Type:
    def __init__(self, shape, dtype):
        self.shape, self.dtype = tuple(int(s) for s in shape), dtype


class _Value:
    def __init__(self, shape, dtype):
        self.type = _Type(shape, dtype)


def _install_stubs(record):
    ir = types.ModuleType("jax.interpreters.mlir.ir")

    class RankedTensorType:
        def __init__(self, t):
            self.shape, self.element_type = list(t.shape), t.dtype

        @staticmethod
        def get(shape, element_type):
            return _Type(shape, element_type)

    class IntegerType:
        get_unsigned = staticmethod(lambda bits: f"uint{bits}")
        get_signless = staticmethod(lambda bits: "bool" if bits == 1 else f"int{bits}")

    class F32Type:
        get = staticmethod(lambda: "float32")

    ir.RankedTensorType, ir.IntegerType, ir.F32Type = RankedTensorType, IntegerType, F32Type
    ir.Value = ir.BlockArgument = _Value
    mlir = types.ModuleType("jax.interpreters.mlir")
    mlir.ir, mlir.LoweringRule, mlir.LoweringRuleContext = ir, object, object
    interpreters = types.ModuleType("jax.interpreters")
    interpreters.mlir = mlir
    jax = types.ModuleType("jax")
    jax.interpreters = interpreters
    jaxlib = types.ModuleType("jaxlib")
    hlo = types.ModuleType("jaxlib.hlo_helpers")  # jaxlib.mhlo_helpers stays absent: the rules fall back, as on a recent jaxlib

    def custom_call(call_target_name, out_types=None, operands=None, backend_config=None, operand_layouts=None,
                    result_layouts=None, **kw):
        assert not kw, f"unrecorded custom_call arguments: {sorted(kw)}"
        record.append(dict(
            target=call_target_name if isinstance(call_target_name, str) else call_target_name.decode(),
            operands=[[list(o.type.shape), o.type.dtype] for o in operands], operand_ids=[id(o) for o in operands],
            results=[[list(t.shape), t.dtype] for t in out_types],
            operand_layouts=[list(x) for x in operand_layouts], result_layouts=[list(x) for x in result_layouts],
            opaque=bytes(backend_config).hex()))
        return [object() for _ in out_types]

    hlo.custom_call = custom_call
    jaxlib.hlo_helpers = hlo
    lib = types.ModuleType("jax.lib")
    registered = []
    lib.xla_client = types.SimpleNamespace(register_custom_call_target=lambda name, value, platform: registered.append((name, value, platform)))
    jax.lib = lib
    sys.modules.update({"jax": jax, "jax.interpreters": interpreters, "jax.interpreters.mlir": mlir,
                        "jax.interpreters.mlir.ir": ir, "jaxlib": jaxlib, "jaxlib.hlo_helpers": hlo, "jax.lib": lib})
    return registered


def _load_lowering(pkg_name, extension_attr, extension_module, path, sub):
    """Import ``path`` (a lowering.py) as ``<pkg_name>.<sub>.lowering`` so that its ``from .. import <extension>``
    finds this repo's drop-in module."""
    pkg = sys.modules.get(pkg_name)
    if pkg is None:
        pkg = types.ModuleType(pkg_name)
        pkg.__path__ = []
        sys.modules[pkg_name] = pkg
    setattr(pkg, extension_attr, extension_module)
    sys.modules[f"{pkg_name}.{extension_attr}"] = extension_module
    subpkg = types.ModuleType(f"{pkg_name}.{sub}")
    subpkg.__path__ = []
    sys.modules[subpkg.__name__] = subpkg
    spec = importlib.util.spec_from_file_location(f"{pkg_name}.{sub}.lowering", path)
    module = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = module
    spec.loader.exec_module(module)
    return module


def _registration_loops(path, namespace):
    """Execute the top-level ``for name, value in <ext>.get_*_registrations().items(): register(...)`` of an impl.py."""
    tree = ast.parse(open(path).read())
    loops = [n for n in tree.body if isinstance(n, ast.For) and "registrations" in ast.unparse(n.iter)]
    assert loops, path
    exec(compile(ast.Module(body=loops, type_ignores=[]), path, "exec"), namespace)


def _ffi_surface(path):
    src = open(path).read()
    getters = {}
    for m in re.finditer(r"pybind11::dict (get_\w+)\(\) \{(.*?)\n\}", src, re.S):
        getters[m.group(1)] = re.findall(r'dict\["(\w+)"\]', m.group(2))
    return {"module": re.search(r"PYBIND11_MODULE\((\w+),", src).group(1), "functions": re.findall(r'm\.def\(\s*"(\w+)"', src),
            "registrations": getters, "capsule_name": re.search(r'pybind11::capsule\([^,]+, "([^"]+)"\)', src).group(1)}


def main():
    import ctypes
    from jaxngp_b200.jaxtcnn import tcnnutils
    from jaxngp_b200.volrendjax import volrendutils_cuda

    record = []
    registered = _install_stubs(record)
    rules = {}
    for sub in ("packbits", "morton3d", "marching", "integrating"):
        mod = _load_lowering("ref_volrendjax", "volrendutils_cuda", volrendutils_cuda,
                             os.path.join(VR, "src", "volrendjax", sub, "lowering.py"), sub)
        rules.update({k: v for k, v in vars(mod).items() if k.endswith("_rule")})
    mod = _load_lowering("ref_jaxtcnn", "tcnnutils", tcnnutils,
                         os.path.join(TC, "src", "jaxtcnn", "hashgrid_tcnn", "lowering.py"), "hashgrid_tcnn")
    rules.update({k: v for k, v in vars(mod).items() if k.endswith("_rule")})
    assert set(rules) == set(CASES), (sorted(rules), sorted(CASES))

    calls = {}
    for name, case in CASES.items():
        before = len(record)
        operands = [_Value(shape, dtype) for shape, dtype in case["arrays"].values()]
        rules[name](None, *operands, **case["statics"])
        assert len(record) == before + 1, name
        entry = record[-1]
        entry["rule"] = name
        by_id = {id(v): n for n, v in zip(case["arrays"], operands)}
        entry["operand_names"] = [by_id[i] for i in entry.pop("operand_ids")]  # which arguments the rule passes on, in order
        entry["rule_arguments"] = {n: [list(shape), dtype] for n, (shape, dtype) in case["arrays"].items()}
        entry["statics"] = case["statics"]
        calls[entry["target"]] = entry

    # the registration loops of the reference's impl.py files, run against the drop-in modules
    xla_client = sys.modules["jax.lib"].xla_client
    for sub in ("packbits", "morton3d", "marching", "integrating"):
        _registration_loops(os.path.join(VR, "src", "volrendjax", sub, "impl.py"),
                            dict(volrendutils_cuda=volrendutils_cuda, xla_client=xla_client))
    _registration_loops(os.path.join(TC, "src", "jaxtcnn", "hashgrid_tcnn", "impl.py"),
                        dict(tcnnutils=tcnnutils, xla_client=xla_client))
    get_name = ctypes.pythonapi.PyCapsule_GetName
    get_name.restype, get_name.argtypes = ctypes.c_char_p, [ctypes.py_object]
    reg = []
    for name, capsule, platform in registered:
        assert get_name(capsule) == b"xla._CUSTOM_CALL_TARGET"
        reg.append([name, platform])
    assert {n for n, _ in reg} == set(calls), (reg, sorted(calls))

    # the custom_vjp of integrate_rays (integrating/impl.py:50-146), unmodified: which residuals and cotangents its bwd rule
    # hands the backward primitive, and which primal argument each returned cotangent is bound to
    import typing
    bound = {}

    class _Prim:
        def __init__(self, name, results):
            self.name, self.results = name, results

        def bind(self, *operands, **statics):
            bound[self.name] = ([o if isinstance(o, str) else "<array>" for o in operands], sorted(statics))
            return tuple(self.results)

    class _CustomVjp:
        def __init__(self, fn):
            self.fn = fn

        def __call__(self, *a, **k):
            return self.fn(*a, **k)

        def defvjp(self, fwd, bwd):
            self.fwd, self.bwd = fwd, bwd

    jax_stub = types.SimpleNamespace(custom_vjp=_CustomVjp, Array=object,
                                     numpy=types.SimpleNamespace(broadcast_to=lambda a, shape: a))
    ns = dict(jax=jax_stub, Tuple=typing.Tuple,
              integrate_rays_p=_Prim("integrate_rays", ["measured_batch_size", "final_rgbds", "final_opacities"]),
              integrate_rays_bwd_p=_Prim("integrate_rays_backward", ["dL_dbgs", "dL_dz_vals", "dL_ddrgbs"]))
    tree = ast.parse(open(os.path.join(VR, "src", "volrendjax", "integrating", "impl.py")).read())
    wanted = {"__integrate_rays", "__fwd_integrate_rays", "__bwd_integrate_rays"}
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in wanted]
    body += [n for n in tree.body if isinstance(n, ast.Expr) and "defvjp" in ast.unparse(n)]
    assert len(body) == 4
    exec(compile(ast.Module(body=body, type_ignores=[]), "integrating/impl.py", "exec"), ns)
    vjp = ns["__integrate_rays"]
    primal_names = ["near_distance", "rays_sample_startidx", "rays_n_samples", "bgs", "dss", "z_vals", "drgbs"]

    class _Named(str):  # a string standing in for an array: only .shape is asked of it (the bgs broadcast)
        shape = (N,)

    primals = {n: _Named(n) for n in primal_names}
    _, aux = vjp.fwd(**primals)
    cotangents = vjp.bwd(aux, ("dL_dmeasured_batch_size", "dL_dfinal_rgbds", "dL_dfinal_opacities"))
    integrate_vjp = {"primal_arguments": primal_names,
                     "backward_operands": bound["integrate_rays_backward"][0], "backward_statics": bound["integrate_rays_backward"][1],
                     "cotangent_bound_to": {n: c for n, c in zip(primal_names, cotangents)}}
    print("integrate_rays bwd binds:", integrate_vjp["cotangent_bound_to"])

    out = {"custom_calls": calls, "registered_by_the_reference_impl": reg, "integrate_rays_vjp": integrate_vjp,
           "ffi": {"volrendjax": _ffi_surface(os.path.join(VR, "lib", "ffi.cc")),
                   "jaxtcnn": _ffi_surface(os.path.join(TC, "lib", "ffi.cc"))}}
    path = os.path.join(ROOT, "tests", "golden", "lowering_reference.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", path, os.path.getsize(path), "bytes;", len(calls), "custom calls,", len(reg), "registrations")
    for t, c in calls.items():
        print(f"  {t:28s} {len(c['operands'])} operands -> {len(c['results'])} results, opaque {len(c['opaque']) // 2} B")


if __name__ == "__main__":
    main()

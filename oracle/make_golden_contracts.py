"""TEST INFRASTRUCTURE ONLY -- the ops' abstract-evaluation contracts from the reference's OWN code.

Runs every ``*_abstract`` rule of volume-rendering-jax ({marching,integrating,packbits,morton3d}/abstract.py) and of
jax-tcnn (hashgrid_tcnn/abstract.py), unmodified, on a table of well-formed and malformed operand signatures (oracle/ref_shim.install_abstract) and records
what it answers: output shapes / dtypes, or the exception class it raises.  tests/test_oracle_golden.py replays the
same table on the host mirror (jaxngp_b200.volrendjax.*).  Writes tests/golden/contracts_reference.json.

    python oracle/make_golden_contracts.py        # needs /root/reference
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

N, S, K, G, CAP, NS = 40, 96, 1, 16, 8, 12  # rays, sample slots, cascades, grid, steps per slot, slots
BITS = K * G ** 3 // 8


def cases():
    """(case name, op, operands as (shape, dtype) in the abstract rule's order, static kwargs)."""
    f32, u32, u8, b, f16 = "float32", "uint32", "uint8", "bool", "float16"
    march = [((N, 3), f32), ((N, 3), f32), ((N,), f32), ((N,), f32), ((N,), f32), ((BITS,), u8)]
    march_st = dict(total_samples=S, diagonal_n_steps=1024, K=K, G=G, bound=1.0, stepsize_portion=0.0)
    minf = [((N, 3), f32), ((N, 3), f32), ((N,), f32), ((N,), f32), ((BITS,), u8), ((1,), u32), ((NS,), b), ((NS,), u32)]
    minf_st = dict(diagonal_n_steps=1024, K=K, G=G, march_steps_cap=CAP, bound=1.0, stepsize_portion=0.0)
    integ = [((N,), u32), ((N,), u32), ((N, 3), f32), ((S,), f32), ((S,), f32), ((S, 4), f32)]
    iinf = [((N, 3), f32), ((N, 4), f32), ((N,), f32), ((NS,), u32), ((NS,), u32), ((NS, CAP), f32), ((NS, CAP), f32), ((NS, CAP, 4), f32)]

    def swap(ops, i, new):
        ops = list(ops)
        ops[i] = new
        return ops

    out = [
        ("march ok", "march_rays_abstract", march, march_st),
        ("march rays_d shape", "march_rays_abstract", swap(march, 1, ((N + 1, 3), f32)), march_st),
        ("march t_ends shape", "march_rays_abstract", swap(march, 3, ((N, 1), f32)), march_st),
        ("march bitfield size", "march_rays_abstract", swap(march, 5, ((BITS - 1,), u8)), march_st),
        ("march bitfield dtype", "march_rays_abstract", swap(march, 5, ((BITS,), "int32")), march_st),
        ("march f16 rays", "march_rays_abstract", swap(swap(march, 0, ((N, 3), f16)), 1, ((N, 3), f16)), march_st),
        ("march zero budget", "march_rays_abstract", march, dict(march_st, total_samples=0)),
        ("march negative portion", "march_rays_abstract", march, dict(march_st, stepsize_portion=-0.1)),
        ("march zero K", "march_rays_abstract", march, dict(march_st, K=0)),
        ("march_inference ok", "march_rays_inference_abstract", minf, minf_st),
        ("march_inference counter shape", "march_rays_inference_abstract", swap(minf, 5, ((2,), u32)), minf_st),
        ("march_inference indices shape", "march_rays_inference_abstract", swap(minf, 7, ((NS + 1,), u32)), minf_st),
        ("march_inference bitfield dtype", "march_rays_inference_abstract", swap(minf, 4, ((BITS,), "int8")), minf_st),
        ("integrate ok", "integrate_rays_abstract", integ, {}),
        ("integrate bgs shape", "integrate_rays_abstract", swap(integ, 2, ((N, 4), f32)), {}),
        ("integrate drgbs shape", "integrate_rays_abstract", swap(integ, 5, ((S, 3), f32)), {}),
        ("integrate f16 drgbs", "integrate_rays_abstract", swap(integ, 5, ((S, 4), f16)), {}),
        ("integrate_inference ok", "integrate_rays_inference_abstract", iinf, {}),
        ("integrate_inference rays_T shape", "integrate_rays_inference_abstract", swap(iinf, 2, ((N, 1), f32)), {}),
        ("integrate_inference drgbs shape", "integrate_rays_inference_abstract", swap(iinf, 7, ((NS, CAP, 3), f32)), {}),
        ("packbits ok", "pack_density_into_bits_abstract", [((64,), f32), ((64,), f32)], {}),
        ("packbits not multiple of 8", "pack_density_into_bits_abstract", [((60,), f32), ((60,), f32)], {}),
        ("packbits f16", "pack_density_into_bits_abstract", [((64,), f16), ((64,), f16)], {}),
        ("packbits rank", "pack_density_into_bits_abstract", [((8, 8), f32), ((8, 8), f32)], {}),
        ("morton3d ok", "morton3d_abstract", [((50, 3), u32)], {}),
        ("morton3d float", "morton3d_abstract", [((50, 3), f32)], {}),
        ("morton3d_invert ok", "morton3d_invert_abstract", [((50,), u32)], {}),
        ("morton3d_invert float", "morton3d_invert_abstract", [((50,), f32)], {}),
    ]
    # jaxtcnn.hashgrid_encode (deps/jax-tcnn/src/jaxtcnn/hashgrid_tcnn/abstract.py): offsets u32[L+1], coords f32[3, n], params f32[rows, F]
    tcnn = [((17,), u32), ((3, 70), f32), ((5000, 2), f32)]
    tcnn_st = dict(L=16, F=2, N_min=16, per_level_scale=1.38)
    out += [
        ("tcnn ok", "hashgrid_encode_abstract", tcnn, tcnn_st),
        ("tcnn 2-D coordinates", "hashgrid_encode_abstract", swap(tcnn, 1, ((2, 70), f32)), tcnn_st),
        ("tcnn offsets length", "hashgrid_encode_abstract", swap(tcnn, 0, ((16,), u32)), tcnn_st),
        ("tcnn params width", "hashgrid_encode_abstract", swap(tcnn, 2, ((5000, 4), f32)), tcnn_st),
        ("tcnn offsets dtype", "hashgrid_encode_abstract", swap(tcnn, 0, ((17,), f32)), tcnn_st),
        ("tcnn f16 coordinates", "hashgrid_encode_abstract", swap(tcnn, 1, ((3, 70), f16)), tcnn_st),
        ("tcnn f16 params", "hashgrid_encode_abstract", swap(tcnn, 2, ((5000, 2), f16)), tcnn_st),
        ("tcnn integer scale", "hashgrid_encode_abstract", tcnn, dict(tcnn_st, per_level_scale=2)),
    ]
    return out


def main():
    from oracle import ref_shim
    rules, SA = ref_shim.install_abstract()
    table = []
    for name, op, operands, static in cases():
        try:
            outs = rules[op](*[SA(s, d) for s, d in operands], **static)
            outs = outs if isinstance(outs, tuple) else (outs,)
            answer = {"ok": [[list(o.shape), str(o.dtype)] for o in outs]}
        except Exception as exc:  # the contract IS the exception class
            answer = {"raises": type(exc).__name__}
        table.append(dict(case=name, op=op, operands=[[list(s), d] for s, d in operands], static=static, answer=answer))
        print(f"{name:36s} {answer if 'raises' in answer else 'ok ' + str(len(answer['ok'])) + ' outputs'}")
    path = os.path.join(ROOT, "tests", "golden", "contracts_reference.json")
    json.dump(table, open(path, "w"), indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()

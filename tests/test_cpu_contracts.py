"""CPU-only checks: the C-ABI library loads and exports every declared symbol, descriptors are
byte-identical to the reference structs, level-table geometry, oracle known-answer tests."""
import ctypes
import os
import re
import struct

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from jaxngp_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    header = open(os.path.join(ROOT, "include", "ngp_b200.h")).read()
    declared = set(re.findall(r"\b(ngp_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.OPS) | set(_lib.STATUS_SYMBOLS)
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert _lib.lib().ngp_b200_abi_version() == 1
    assert _lib.lib().ngp_b200_last_status() == 0


def test_no_cpu_fallback():
    import torch
    from jaxngp_b200 import _lib, volrendjax as V
    with pytest.raises(_lib.NgpError):
        V.morton3d(torch.zeros(4, 3, dtype=torch.int32))  # CPU tensor: refused, never computed on the host


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "jaxngp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle", "").replace("CPU oracle", ""), os.path.join(dirpath, f)


def test_descriptor_wire_format():
    from jaxngp_b200 import descriptors as D
    # sizes of the reference structs (volrend.h:23-118, tcnnutils.h:11-29)
    assert len(D.make_packbits_descriptor(5)) == 4
    assert len(D.make_morton3d_descriptor(5)) == 4
    assert len(D.make_marching_descriptor(1, 2, 3, 4, 5, 1.0, 0.0)) == 28
    assert len(D.make_marching_inference_descriptor(1, 2, 3, 4, 5, 6, 1.0, 0.0)) == 32
    assert len(D.make_integrating_descriptor(1, 2)) == 8
    assert len(D.make_integrating_backward_descriptor(1, 2, 0.3)) == 12
    assert len(D.make_integrating_inference_descriptor(1, 2, 3)) == 12
    assert len(D.make_hashgrid_descriptor(1, 16, 2, 16, 1.38)) == 20
    assert D.make_integrate_loss_descriptor(3, 9, 0.3, 0.1) == struct.pack("<2I2f", 3, 9, 0.3, 0.1)  # NgpIntegrateLossDescriptor
    assert D.make_marching_descriptor(7, 9, 1024, 1, 128, 1.0, 0.5) == struct.pack("<IIIIIff", 7, 9, 1024, 1, 128, 1.0, 0.5)
    with pytest.raises(RuntimeError):
        D.make_marching_descriptor(1, 1, 1, 0, 1, 1.0, 0.0)  # ffi.cc:79-81
    with pytest.raises(RuntimeError):
        D.make_packbits_descriptor(0)  # ffi.cc:57-59
    a1 = D.make_hashgrid_a1_descriptor(10, 3, 16, 2, 1 << 19, 0, 1.0, [0] * 5 + [1] * 11, [1.0] * 16, [2] * 16, list(range(17)))
    assert len(a1) == 32 + 4 * (32 + 32 + 33) + 4  # + rows_per_group


def test_level_table_matches_reference_defaults():
    from jaxngp_b200 import encoders as E
    lt = E.make_level_table(16, 2 ** 19, 2, 16, 2048, 3)
    assert lt.rows == 6098120 and sum(lt.hashed) == 11 and lt.res[0] == 16 and lt.res[-1] == 2049
    assert E.make_level_table(16, 2 ** 19, 2, 16, 2048, 3, align=1).rows == 6098108
    lt2 = E.make_level_table(16, 2 ** 19, 2, 16, 2 ** 19, 2)
    assert lt2.rows == 5592320 and abs(lt2.b - 2.0) < 1e-12  # imagefit shape (C1)


def test_oracle_known_answers(oracle):
    # Morton (marching.cu:52-77)
    assert oracle.morton3d(np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1023, 1023, 1023]], np.uint32)).tolist() == [1, 2, 4, 0x3FFFFFFF]
    rng = np.random.default_rng(0)
    xyz = rng.integers(0, 1024, (1000, 3), dtype=np.uint32)
    assert np.array_equal(oracle.morton3d_invert(oracle.morton3d(xyz)), xyz)
    # packbits LSB first (packbits.cu:28-33)
    assert oracle.packbits(0.5, np.array([1, 0, 0, 0, 0, 0, 0, 1], np.float32))[1].tolist() == [0x81]
    # uniform cube a la make_test_cube (models/nerfs.py:477-503): all-ones grid, constant sigma
    G, steps = 128, 1024
    bits = np.full(G ** 3 // 8, 0xFF, np.uint8)
    o = np.array([[0.0, 0.0, -2.0]], np.float32)
    d = np.array([[0.0, 0.0, 1.0]], np.float32)
    ts, te = np.array([1.0], np.float32), np.array([3.0], np.float32)
    mb, valid, rn, rs, idcs, xyzs, dirs, dss, zs = oracle.march_rays(4096, steps, 1, G, 1.0, 0.0, o, d, ts, te, 0.0, bits)
    ds = np.float32(2 * np.sqrt(3) / steps)
    n_expected = int(np.ceil(2.0 / float(ds)))
    assert abs(int(rn[0]) - n_expected) <= 1 and valid[0] and mb == rn[0]
    assert np.allclose(dss[: rn[0]], ds) and np.allclose(np.diff(zs[: rn[0]]), ds, atol=1e-6)
    sigma = 0.7
    drgbs = np.zeros((4096, 4), np.float32)
    drgbs[:, 0] = sigma
    drgbs[:, 1:] = 0.25
    mbs, rgbd, opac = oracle.integrate_rays(0.3, rs, rn, np.zeros(3, np.float32), dss, zs, drgbs)
    assert np.isclose(opac[0], 1 - np.exp(-sigma * float(ds) * int(rn[0])), atol=1e-5)
    assert np.allclose(rgbd[0, :3], 0.25 * opac[0], atol=1e-5) and mbs == rn[0]


def test_oracle_hashgrid_two_restatements_agree(oracle):
    from oracle import hashgrid_np as H
    for dim, T, N_max in ((3, 2 ** 19, 2048), (2, 2 ** 14, 2 ** 12)):
        lv = H.level_table(16, T, 2, 16, N_max, dim)
        rng = np.random.default_rng(3)
        pts = rng.uniform(-1, 1, (3000, dim)).astype(np.float32)
        tab = rng.uniform(-1, 1, (int(lv["offsets"][-1]), 2)).astype(np.float32)
        assert np.allclose(oracle.hashgrid_encode(lv, pts, 1.0, tab), H.encode(lv, pts, 1.0, tab), atol=1e-6)
        d = rng.normal(size=(3000, 32)).astype(np.float32)
        assert np.allclose(oracle.hashgrid_backward(lv, pts, 1.0, d, 2), H.backward(lv, pts, 1.0, d, 2), atol=1e-9)
    # Q1: dense levels spill past their own rows because the modulus is T (encoders.py:187)
    lv = H.level_table(16, 2 ** 19, 2, 16, 2048, 3)
    idx, _ = H.indices_and_weights(lv, np.array([[0.99, 0.99, 0.99]], np.float32), 1.0)
    assert idx[0].max() >= lv["offsets"][1]
    # grid vertex: interpolation returns the stored feature exactly
    tab = np.arange(int(lv["offsets"][-1]) * 2, dtype=np.float32).reshape(-1, 2)
    p = np.array([[(3 - 0.5) / 15 * 2 - 1, (4 - 0.5) / 15 * 2 - 1, (5 - 0.5) / 15 * 2 - 1]], np.float32)  # level-0 vertex (3,4,5)
    e = oracle.hashgrid_encode(lv, p, 1.0, tab)
    assert np.allclose(e[0, :2], tab[3 + 4 * 16 + 5 * 256], rtol=1e-4)


def test_adam_exchange_descriptor_and_argument_checks(monkeypatch, built_lib):
    """ngp_adam_step_exchange (csrc/exchange.cu): wire format of its descriptor, and the argument checks that run
    before anything is enqueued (they need no GPU: a refused call returns before the launch)."""
    from jaxngp_b200 import _lib, descriptors as D, exchange as X

    class Adam(ctypes.Structure):
        _fields_ = [("n", ctypes.c_uint64), ("decay_begin", ctypes.c_uint64), ("lr_init", ctypes.c_float),
                    ("lr_end", ctypes.c_float), ("decay_rate", ctypes.c_float), ("transition_steps", ctypes.c_uint32),
                    ("transition_begin", ctypes.c_uint32), ("staircase", ctypes.c_uint32), ("b1", ctypes.c_float),
                    ("b2", ctypes.c_float), ("eps", ctypes.c_float), ("eps_root", ctypes.c_float),
                    ("weight_decay", ctypes.c_float), ("grad_scale", ctypes.c_float)]

    class Exchange(ctypes.Structure):
        _fields_ = [("adam", Adam), ("shard_begin", ctypes.c_uint64), ("rank", ctypes.c_uint32), ("world", ctypes.c_uint32),
                    ("use_multimem", ctypes.c_uint32), ("n_blocks", ctypes.c_uint32), ("signal_base", ctypes.c_uint32),
                    ("timeout_ms", ctypes.c_uint32)]

    adam = D.make_adam_descriptor(n=1024, decay_begin=512, lr_init=1e-2, lr_end=1e-4, decay_rate=1 / 3, transition_steps=10_000,
                                  transition_begin=10_000, staircase=True, b1=0.9, b2=0.99, eps=1e-15, eps_root=1e-15,
                                  weight_decay=1e-6, grad_scale=0.125)
    raw = D.make_adam_exchange_descriptor(adam, shard_begin=4096, rank=3, world=8, use_multimem=True, n_blocks=96,
                                          signal_base=X.SIGNAL_BASE)
    assert len(raw) == ctypes.sizeof(Exchange) == 96 and len(adam) == ctypes.sizeof(Adam) == 64
    d = Exchange.from_buffer_copy(raw)
    assert (d.adam.n, d.adam.decay_begin, d.shard_begin, d.rank, d.world, d.use_multimem, d.n_blocks, d.signal_base) == \
        (1024, 512, 4096, 3, 8, 1, 96, 1024)
    assert d.adam.grad_scale == 0.125 and d.adam.staircase == 1

    L = built_lib
    bufs = (ctypes.c_void_p * 8)()

    def status(desc):
        L.ngp_b200_clear_error()
        L.ngp_adam_step_exchange(None, bufs, desc, len(desc))
        st = L.ngp_b200_last_status()
        L.ngp_b200_clear_error()
        return st

    assert status(raw[:-4]) == -1                                                                  # descriptor size
    assert status(D.make_adam_exchange_descriptor(adam, 4098, 3, 8, False, 96, 1024)) == -2         # shard not float4-aligned
    assert status(D.make_adam_exchange_descriptor(adam, 4096, 8, 8, False, 96, 1024)) == -2         # rank outside the world
    assert status(D.make_adam_exchange_descriptor(adam, 4096, 0, 9, False, 96, 1024)) == -2         # more than 8 ranks
    assert status(D.make_adam_exchange_descriptor(adam, 4096, 0, 8, False, 149, 1024)) == -2        # more CTAs than SMs
    assert status(raw) == -2                                                                        # multimem without multicast bases

    # host side: the mode switch and the CTA budget of the signal pads
    monkeypatch.delenv("NGP_B200_EXCHANGE", raising=False)
    assert X.requested_mode() == "auto"  # the fused kernel, NCCL if peers cannot be mapped
    monkeypatch.setenv("NGP_B200_EXCHANGE", "peer")
    assert X.requested_mode() == "peer"
    monkeypatch.setenv("NGP_B200_EXCHANGE", "gloo")
    with pytest.raises(ValueError):
        X.requested_mode()
    assert X.blocks_for(9216, 8) == 64 and X.blocks_for(9216, 2) == 64 and X.blocks_for(9216, 2, want=148) == 148 and X.blocks_for(4096 + 4 * 8 * 5, 8) == 5
    with pytest.raises(_lib.NgpError):
        X.blocks_for(4096, 8)
    with pytest.raises(_lib.NgpError):  # no process group on this host: the peer exchange refuses, nothing falls back
        X.PeerExchange(1 << 12, 0, 2, "cpu")


def test_bench_roof_is_chosen_by_arithmetic_intensity():
    """bench.pick_roof: a kernel that contracts and streams is held to the lower roof at its FLOP/byte.  The MLP backward alone
    (56,448 FLOP over 284 B per sample) sits under the tensor roof, the step's backward with the table scatter fused in
    (1192 B per sample + the table zero-fill) under the HBM roof; `frac` is always achieved / peak of the roof named."""
    import bench
    n, tpeak, hbm = 1 << 18, 842.8, 6543.7
    a = bench.pick_roof(n * 56448, n * 284, 0.132, tpeak, hbm)
    assert a["bound"] == "tensor" and a["unit"] == "TFLOP/s" and abs(a["achieved"] - n * 56448 / 0.132e-3 / 1e12) < 0.01
    assert abs(a["frac"] - a["achieved"] / tpeak) < 1e-3 and a["frac"] == a["frac_of_tensor_peak"]
    b = bench.pick_roof(n * 56448, n * 1192 + 48_784_960, 0.204, tpeak, hbm)
    assert b["bound"] == "hbm" and b["unit"] == "GB/s" and abs(b["frac"] - b["achieved"] / hbm) < 1e-3
    assert b["arithmetic_intensity_flop_per_byte"] < b["machine_balance_flop_per_byte"] < a["arithmetic_intensity_flop_per_byte"]
    for r in (a, b):
        assert 0 < r["frac_of_tensor_peak"] < 1 and 0 < r["frac_of_hbm_peak"] < 1

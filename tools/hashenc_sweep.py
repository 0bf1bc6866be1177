"""C4 sweep of the hash-grid encoder (2^22 points, L=16, F=2) over the table size T and the levels-per-pass policy of
csrc/hashgrid.cu ("level-major passes"): time forward / backward for each forced policy, check every policy against the
single-pass kernels (forward bit-identical, backward to atomic-order tolerance), print one JSON line per T.
Usage: python tools/hashenc_sweep.py [log2_T ...]      (policies forced through NGP_B200_HG_LPG / NGP_B200_HG_BWD_LPG)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from jaxngp_b200 import encoders as E


ONCE = os.environ.get("HASHENC_SWEEP_ONCE") == "1"  # under ncu: one launch per configuration, no timing loops


def timed(fn, iters=8):
    if ONCE:
        fn()
        torch.cuda.synchronize()
        return 0.0
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    dev = torch.device("cuda", 0)
    n = 1 << 22
    g = torch.Generator(device=dev)
    pos = torch.rand(n, 3, device=dev, generator=g.manual_seed(42)) * 2 - 1
    d_enc = torch.randn(n, 32, device=dev, generator=g.manual_seed(44))
    for lt2 in [int(a) for a in sys.argv[1:]] or (19, 20, 21, 22, 23, 24):
        lt = E.make_level_table(16, 2 ** lt2, 2, 16, 2048, 3)
        table = (torch.rand(lt.rows, 2, device=dev, generator=g.manual_seed(43)) * 2 - 1)
        grad = torch.empty_like(table)
        row = {"log2_T": lt2, "table_mb": round(table.numel() * 4 / 2 ** 20, 1), "fwd_ms": {}, "bwd_ms": {}}
        os.environ["NGP_B200_HG_LPG"] = os.environ["NGP_B200_HG_BWD_LPG"] = "16"
        ref_enc = E.hashgrid_forward(lt, pos, 1.0, table)
        ref_grad = E.hashgrid_backward(lt, pos, 1.0, d_enc).clone()
        for lpg in ((16,) if ONCE else (16, 8, 4, 2, 1)):
            os.environ["NGP_B200_HG_LPG"] = os.environ["NGP_B200_HG_BWD_LPG"] = str(lpg)
            enc = E.hashgrid_forward(lt, pos, 1.0, table)
            assert torch.equal(enc, ref_enc), (lt2, lpg)
            E.hashgrid_backward(lt, pos, 1.0, d_enc, out=grad)
            err = float((grad - ref_grad).abs().max() / ref_grad.abs().max())
            assert err < 1e-4, (lt2, lpg, err)
            row["fwd_ms"][lpg] = round(timed(lambda: E.hashgrid_forward(lt, pos, 1.0, table)), 3)
            row["bwd_ms"][lpg] = round(timed(lambda: E.hashgrid_backward(lt, pos, 1.0, d_enc, out=grad)), 3)
        del os.environ["NGP_B200_HG_LPG"], os.environ["NGP_B200_HG_BWD_LPG"]
        row["auto_fwd_ms"] = round(timed(lambda: E.hashgrid_forward(lt, pos, 1.0, table)), 3)
        row["auto_bwd_ms"] = round(timed(lambda: E.hashgrid_backward(lt, pos, 1.0, d_enc, out=grad)), 3)
        alg = n * 1164 * 2 + table.numel() * 4
        row["auto_fwd_bwd_gbs"] = 0.0 if ONCE else round(alg / (row["auto_fwd_ms"] + row["auto_bwd_ms"]) / 1e6, 1)
        print(json.dumps(row), flush=True)
        del table, grad, ref_enc, ref_grad


if __name__ == "__main__":
    main()

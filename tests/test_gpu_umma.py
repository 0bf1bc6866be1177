"""tcgen05 / TMEM operand formats (csrc/umma.cuh): the three GEMM shapes of the tensor-core MLP against torch."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _tf32(x):
    """round-to-nearest (ties away) to 10 mantissa bits, like cvt.rna.tf32.f32"""
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def test_umma_selftest_matches_torch():
    from jaxngp_b200 import _lib
    g = torch.Generator(device=DEV).manual_seed(5)
    A = torch.randn(128, 32, device=DEV, generator=g)
    W = torch.randn(32, 64, device=DEV, generator=g)
    G = torch.randn(128, 64, device=DEV, generator=g)
    D1, D2, D3 = torch.empty(128, 64, device=DEV), torch.empty(128, 32, device=DEV), torch.empty(64, 32, device=DEV)
    _lib.call("ngp_umma_selftest", [A, W, G, D1, D2, D3], b"")
    torch.cuda.synchronize()
    a, w, gg = _tf32(A).double(), _tf32(W).double(), _tf32(G).double()
    for name, got, ref in (("A.W", D1, a @ w), ("G.W^T", D2, gg @ w.T), ("G^T.A", D3, gg.T @ a)):
        err = (got.double() - ref).abs().max().item()
        assert err < 2e-4 * ref.abs().max().item() + 1e-5, (name, err)

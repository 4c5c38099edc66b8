"""Tensor-core path (tcgen05 kind::tf32, accumulator in TMEM) against fp64 on the same seeded inputs.
n_split = 1 is plain TF32 (10-bit mantissa, ~1e-3); n_split = 3 is the 3xTF32 split the MLP kernel uses and must meet the
1e-4 bar of BASELINE.json's north_star with a wide margin (expected ~1e-6)."""
import pytest
import torch

from tests.util import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,K", [(16, 8), (16, 32), (32, 64), (64, 128), (128, 64), (128, 96), (240, 64), (256, 32)])
def test_tc_selftest_gemm(cuda, N, K):
    from diffusion_edf_b200 import ops
    g = torch.Generator().manual_seed(N * 1000 + K)
    A = torch.randn(128, K, generator=g)
    B = torch.randn(N, K, generator=g)
    ref = A.double() @ B.double().t()
    d3 = ops.tc_selftest(A.to(cuda), B.to(cuda), 3)
    e3 = rel_err(d3, ref)
    assert e3 < 5e-6, f"3xTF32 N={N} K={K}: {e3:.3e}"
    d1 = ops.tc_selftest(A.to(cuda), B.to(cuda), 1)
    e1 = rel_err(d1, ref)
    assert 1e-6 < e1 < 5e-3, f"plain TF32 N={N} K={K}: {e1:.3e}"      # it really ran at tf32 precision

"""Out = P^T Q over the long dimension (lob_tn_matmul): the register-tiled fp32 kernel for the preconditioner's Q^T r
shape (csrc/tn_skinny.cu) and the generic kernel it falls back to, against an fp64 product of the same inputs.
Reference call site: operators/added_diag_linear_operator.py:137 (`qqt_term = q @ (q.mT @ tensor)`)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from linear_operator_b200 import _kernels  # noqa: E402

DEV = "cuda:0"


@pytest.mark.parametrize(
    "B,N,I,J",
    [(3, 5000, 100, 33), (2, 1000, 8, 1), (2, 777, 128, 48), (1, 300, 52, 17), (2, 4096, 100, 32), (5, 260, 12, 5),
     (2, 1000, 130, 33), (2, 1000, 50, 33), (1, 200, 100, 33), (2, 100003, 16, 33), (3, 3001, 24, 11), (1, 999, 32, 48)],
)
def test_tn_matmul_matches_fp64(B, N, I, J):
    g = torch.Generator(device=DEV).manual_seed(B * 1000 + N + I + J)
    P = torch.randn(B, N, I, device=DEV, generator=g)
    Q = torch.randn(B, N, J, device=DEV, generator=g)
    ref = P.double().mT @ Q.double()
    out = _kernels.tn_matmul(P, Q)
    # fp32 FMA accumulation, round to nearest: error ~ eps * sqrt(N) relative to the magnitude of the terms
    bound = 4e-7 * N**0.5 * (P.double().abs().mT @ Q.double().abs())
    assert ((out.double() - ref).abs() <= bound + 1e-30).all()
    # the generic CUDA-core kernel (what shapes outside the skinny kernel's range take) obeys the same bound
    out2 = _kernels.tn_matmul(P.double(), Q.double())
    assert ((out2 - ref).abs() <= 1e-12 * (P.double().abs().mT @ Q.double().abs()) + 1e-30).all()
    # repeated launches are bit-identical (fixed reduction order)
    assert torch.equal(out, _kernels.tn_matmul(P, Q))


def test_tn_matmul_is_unbiased_on_positive_data():
    """Why Q^T r is not on the tensor cores: no systematic loss of magnitude on all-positive data."""
    g = torch.Generator(device=DEV).manual_seed(1)
    P = torch.rand(2, 5000, 100, device=DEV, generator=g)
    Q = torch.rand(2, 5000, 33, device=DEV, generator=g)
    ref = P.double().mT @ Q.double()
    out = _kernels.tn_matmul(P, Q).double()
    rel = (out - ref) / ref
    assert rel.abs().max().item() < 3e-6
    assert abs(rel.mean().item()) < 3e-7

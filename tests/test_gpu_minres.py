"""shifted MINRES (utils/minres.py) and contour-integral quadrature (utils/contour_integral_quad.py)
against the reference's own outputs and against dense algebra."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import linear_operator_b200 as lo  # noqa: E402
from linear_operator_b200 import settings  # noqa: E402
from linear_operator_b200.operators import AddedDiagLinearOperator, DenseLinearOperator, DiagLinearOperator  # noqa: E402
from linear_operator_b200.utils import contour_integral_quad, minres  # noqa: E402
from test_gpu_parity import DEV, F32_RTOL, F64_RTOL, check, cu, npy  # noqa: E402


def test_minres_shifted_solves_vs_reference(golden):
    g = golden("minres_f64")
    A, rhs, sh = cu(g["A"]), cu(g["rhs"]), cu(g["shifts"])
    sol = minres(A, rhs, shifts=sh, max_iter=100)
    assert sol.shape == g["solve"].shape
    check(npy(sol), g["solve"], F64_RTOL)
    assert float(sol[..., 1].abs().max()) == 0.0  # zero right-hand-side column
    minv = cu(g["minv"]).unsqueeze(-1)
    check(npy(minres(lambda v: lo._kernels.dense_matmul(A, v), rhs, shifts=sh, max_iter=100,
                     preconditioner=lambda v: v * minv)), g["solve_precond"], F64_RTOL)
    check(npy(minres(A, rhs, shifts=-sh - 0.1, value=-1, max_iter=100)), g["solve_neg"], F64_RTOL)
    vec = minres(A[0], rhs[0, :, 0], max_iter=100)
    assert vec.shape == g["solve_vec"].shape
    check(npy(vec), g["solve_vec"], F64_RTOL)
    # and against dense algebra: (K + s I) x = b
    eye = torch.eye(50, device=DEV, dtype=torch.float64)
    for q in range(3):
        want = torch.linalg.solve(A + sh[q] * eye, rhs)
        check(npy(sol[q][..., [0, 2]]), npy(want[..., [0, 2]]), 1e-6)  # minres_tolerance = 1e-4 on the update term


def test_minres_fp32_vs_reference(golden):
    g = golden("minres_f32")
    sol = minres(cu(g["A"]), cu(g["rhs"]), shifts=cu(g["shifts"]), max_iter=100)
    check(npy(sol), g["solve"], F32_RTOL)


def test_contour_integral_quad_vs_reference(golden):
    g = golden("ciq_f64")
    op = AddedDiagLinearOperator(DenseLinearOperator(cu(g["A"])), DiagLinearOperator(cu(g["d"])))
    rhs = cu(g["rhs"])
    dense = cu(g["A"]) + torch.diag_embed(cu(g["d"]))
    evals, evecs = torch.linalg.eigh(dense)
    for tag, inv, power in (("inv", True, -0.5), ("sqrt", False, 0.5)):
        solves, weights, no_shift, shifts = contour_integral_quad(op, rhs, inverse=inv)
        assert solves.shape == g[f"solves_{tag}"].shape and weights.shape == g[f"weights_{tag}"].shape
        check(npy(shifts), g[f"shifts_{tag}"], 1e-9)      # Lanczos eigenvalue bounds -> elliptic functions
        check(npy(weights), g[f"weights_{tag}"], 1e-9)
        res = (solves * weights).sum(0)
        check(npy(res), g[f"result_{tag}"], 1e-8)
        check(npy(no_shift), g[f"no_shift_{tag}"], 1e-8)
        exact = (evecs * evals.pow(power).unsqueeze(-2)) @ evecs.mT @ rhs
        check(npy(res), npy(exact), 1e-4)  # quadrature error with 15 points


def test_ciq_samples_setting_routes_zero_mean_mvn_samples():
    gen = torch.Generator(device=DEV).manual_seed(3)
    W = torch.randn(60, 60, device=DEV, dtype=torch.float64, generator=gen)
    op = AddedDiagLinearOperator(DenseLinearOperator(W @ W.mT / 60), DiagLinearOperator(torch.full((60,), 0.5, device=DEV,
                                                                                                   dtype=torch.float64)))
    with settings.ciq_samples(True):
        torch.manual_seed(0)
        s = op.zero_mean_mvn_samples(4)
    assert s.shape == (4, 60) and torch.isfinite(s).all()
    torch.manual_seed(0)
    eps = torch.randn(60, 4, device=DEV, dtype=torch.float64)
    dense = W @ W.mT / 60 + 0.5 * torch.eye(60, device=DEV, dtype=torch.float64)
    evals, evecs = torch.linalg.eigh(dense)
    want = ((evecs * evals.sqrt()) @ evecs.mT @ eps).mT
    check(npy(s), npy(want), 1e-4)

"""GPU parity of the backward pass against gradients produced by the reference's own autograd
(fixtures of tests/golden/make_golden_backward.py and make_golden_round2.py) and against the oracle at a larger size.
fp64 bar 1e-10 relative where the forward is run to tight tolerance; gradients inherit the CG tolerance otherwise."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import linear_operator_b200 as lo  # noqa: E402
from linear_operator_b200 import _kernels, settings  # noqa: E402
from linear_operator_b200.operators import (  # noqa: E402
    AddedDiagLinearOperator,
    ConstantDiagLinearOperator,
    DenseLinearOperator,
    DiagLinearOperator,
    ToeplitzLinearOperator,
)
from oracle import krylov_oracle as ko  # noqa: E402
from test_gpu_parity import DEV, F32_RTOL, Injected, check, cu, npy  # noqa: E402

GRAD_RTOL = 1e-9  # gradients are products of two CG solves run to tolerance 1e-10


def leaf(a):
    return cu(a).requires_grad_(True)


@pytest.mark.parametrize("name", ["backward_dense_noprecond_f64", "backward_dense_precond_f64",
                                  "backward_dense_const_full_f64", "backward_dense_const_constop_f64"])
def test_inv_quad_logdet_backward_vs_reference(golden, name):
    """InvQuadLogdet.backward + Dense / Diag _bilinear_derivative + PivotedCholesky.backward + d logdet_P."""
    g = golden(name)
    A, d, rhs = leaf(g["A"]), leaf(g["d"]), leaf(g["rhs"])
    precond = "noprecond" not in name
    diag_op = ConstantDiagLinearOperator(d, diag_shape=A.shape[-1]) if "constop" in name else DiagLinearOperator(d)
    op = Injected(DenseLinearOperator(A), diag_op)
    op.probes = cu(g["probes"])
    with settings.max_cholesky_size(0), settings.min_preconditioning_size(4 if precond else 10**6), \
            settings.max_preconditioner_size(int(g["rank"])), settings.cg_tolerance(1e-10), \
            settings.max_cg_iterations(400):
        iq, ld = op.inv_quad_logdet(rhs, logdet=True)
        assert iq.requires_grad and ld.requires_grad
        (cu(g["w_iq"]) * iq + cu(g["w_ld"]) * ld).sum().backward()
    check(npy(iq), g["inv_quad"], 1e-10)
    check(npy(ld), g["logdet"], 1e-10)
    check(npy(A.grad), g["grad_A"], GRAD_RTOL)
    check(npy(d.grad), g["grad_d"], GRAD_RTOL)
    check(npy(rhs.grad), g["grad_rhs"], GRAD_RTOL)


@pytest.mark.parametrize("name", ["backward_solve_f64", "backward_solve_left_f64"])
def test_solve_backward_vs_reference(golden, name):
    g = golden(name)
    A, d, rhs, lhs = leaf(g["A"]), leaf(g["d"]), leaf(g["rhs"]), leaf(g["lhs"])
    has_left = "left" in name
    op = AddedDiagLinearOperator(DenseLinearOperator(A), DiagLinearOperator(d))
    with settings.max_cholesky_size(0), settings.min_preconditioning_size(4), \
            settings.max_preconditioner_size(int(g["rank"])), settings.cg_tolerance(float(g["tol"])), \
            settings.max_cg_iterations(300):
        res = op.solve(rhs, lhs) if has_left else op.solve(rhs)
        (res * cu(g["w"])).sum().backward()
    check(npy(res), g["res"], 1e-10)
    check(npy(A.grad), g["grad_A"], GRAD_RTOL)
    check(npy(d.grad), g["grad_d"], GRAD_RTOL)
    check(npy(rhs.grad), g["grad_rhs"], GRAD_RTOL)
    if has_left:
        check(npy(lhs.grad), g["grad_lhs"], GRAD_RTOL)


def test_inv_quad_backward_vs_reference(golden):
    g = golden("backward_inv_quad_f64")
    A, d, rhs = leaf(g["A"]), leaf(g["d"]), leaf(g["rhs"])
    op = AddedDiagLinearOperator(DenseLinearOperator(A), DiagLinearOperator(d))
    with settings.max_cholesky_size(0), settings.min_preconditioning_size(4), \
            settings.max_preconditioner_size(int(g["rank"])), settings.cg_tolerance(float(g["tol"])), \
            settings.max_cg_iterations(300):
        res = op.inv_quad(rhs, reduce_inv_quad=False)
        (res * cu(g["w"])).sum().backward()
    check(npy(res), g["res"], 1e-10)
    check(npy(A.grad), g["grad_A"], GRAD_RTOL)
    check(npy(d.grad), g["grad_d"], GRAD_RTOL)
    check(npy(rhs.grad), g["grad_rhs"], GRAD_RTOL)


def test_toeplitz_inv_quad_logdet_backward_vs_reference(golden):
    """Column gradient through sym_toeplitz_derivative_quadratic_form (utils/toeplitz.py:164-204)."""
    g = golden("backward_toeplitz_f64")
    col, d, rhs = leaf(g["col"]), leaf(g["d"]), leaf(g["rhs"])
    op = Injected(ToeplitzLinearOperator(col), DiagLinearOperator(d))
    op.probes = cu(g["probes"])
    with settings.max_cholesky_size(0), settings.max_preconditioner_size(0):
        iq, ld = op.inv_quad_logdet(rhs, logdet=True)
        (cu(g["w_iq"]) * iq + cu(g["w_ld"]) * ld).sum().backward()
    check(npy(iq), g["inv_quad"], 1e-10)
    check(npy(ld), g["logdet"], 1e-10)
    check(npy(col.grad), g["grad_col"], 1e-10)
    check(npy(d.grad), g["grad_d"], 1e-10)
    check(npy(rhs.grad), g["grad_rhs"], 1e-10)


def test_bilinear_kernels_against_einsum():
    gen = torch.Generator(device=DEV).manual_seed(4)
    for dtype, tol in ((torch.float32, 1e-5), (torch.float64, 1e-13)):
        left = torch.randn(3, 301, 33, device=DEV, generator=gen, dtype=dtype)
        right = torch.randn(3, 157, 33, device=DEV, generator=gen, dtype=dtype)
        w = torch.randn(3, 33, device=DEV, generator=gen, dtype=dtype)
        want = torch.einsum("bic,bjc,bc->bij", left.double(), right.double(), w.double())
        check(npy(_kernels.bilinear_dense(left, right, w)), npy(want), tol)
        acc = torch.ones(3, 301, 157, device=DEV, dtype=dtype)
        _kernels.bilinear_dense(left, right, None, out=acc)
        check(npy(acc), npy(torch.einsum("bic,bjc->bij", left.double(), right.double()) + 1), tol)
        wantd = torch.einsum("bic,bic,bc->bi", left.double(), left.double(), w.double())
        check(npy(_kernels.bilinear_diag(left, left, w)), npy(wantd), tol)
        C = torch.tril(torch.randn(4, 37, 37, device=DEV, generator=gen, dtype=dtype)) + 6 * torch.eye(37, device=DEV, dtype=dtype)
        check(npy(_kernels.tri_inverse(C)), npy(torch.linalg.inv(C.double())), tol * 10)


def test_bilinear_dense_tensor_core_path_against_einsum():
    """N, M >= 1024 in fp32: the gradient G = L diag(w) R^T goes through csrc/gemm3x.cu (3xTF32, K = 33 padded to 36);
    ragged against the 128-wide tiles."""
    gen = torch.Generator(device=DEV).manual_seed(5)
    left = torch.randn(2, 1100, 33, device=DEV, generator=gen)
    right = torch.randn(2, 1028, 33, device=DEV, generator=gen)
    w = torch.randn(2, 33, device=DEV, generator=gen)
    want = torch.einsum("bic,bjc,bc->bij", left.double(), right.double(), w.double())
    check(npy(_kernels.bilinear_dense(left, right, w)), npy(want), 1e-5)
    want = torch.einsum("bic,bjc->bij", left.double(), right.double())
    check(npy(_kernels.bilinear_dense(left, right)), npy(want), 1e-5)


def test_inv_quad_logdet_backward_mid_size_fp32_vs_oracle():
    """N = 700 (ragged against the 128-wide gradient tiles), batch 2, fp32, rank-12 preconditioner: gradients against
    the oracle's restatement of the reference backward on identical inputs and probes."""
    gen = torch.Generator(device=DEV).manual_seed(21)
    B, N, S = 2, 700, 8
    W = torch.randn(B, N, 64, device=DEV, generator=gen)
    sc = torch.logspace(0, -1.5, 64, device=DEV)
    A = ((W * sc / sc.norm()) @ (W * sc / sc.norm()).mT).requires_grad_(True)
    d = (0.4 + 0.2 * torch.rand(B, N, device=DEV, generator=gen)).requires_grad_(True)
    rhs = torch.randn(B, N, 1, device=DEV, generator=gen).requires_grad_(True)
    probes = torch.randn(B, N, S, device=DEV, generator=gen)
    probes = probes / probes.norm(dim=-2, keepdim=True)
    w_iq = torch.tensor([0.7, -1.3], device=DEV)
    w_ld = torch.tensor([1.1, 0.4], device=DEV)
    op = Injected(DenseLinearOperator(A), DiagLinearOperator(d))
    op.probes = probes
    with settings.max_cholesky_size(0), settings.min_preconditioning_size(100), settings.max_preconditioner_size(12), \
            settings.cg_tolerance(1e-6), settings.max_cg_iterations(200):
        iq, ld = op.inv_quad_logdet(rhs, logdet=True)
        (w_iq * iq + w_ld * ld).sum().backward()
    f64 = lambda t: npy(t).astype(np.float64)  # noqa: E731
    gA, gd, gr = ko.dense_added_diag_inv_quad_logdet_backward(
        f64(A), f64(d), f64(rhs), f64(probes), f64(w_iq), f64(w_ld), precond_rank=12, min_precond_size=100,
        tolerance=1e-6, max_iter=200)
    check(npy(A.grad), gA, 10 * F32_RTOL)  # products of two fp32 solves (each <= 1e-4)
    check(npy(d.grad), gd, 10 * F32_RTOL)
    check(npy(rhs.grad), gr, F32_RTOL)


def test_gradients_are_not_silently_dropped():
    """outputs of operators / right-hand sides that require grad must carry a grad_fn, and an operator
    class without a closed-form derivative must fail loudly instead of returning no gradient."""
    from linear_operator_b200.operators import LinearOperator

    gen = torch.Generator(device=DEV).manual_seed(2)
    A = torch.randn(50, 50, device=DEV, dtype=torch.float64, generator=gen)
    A = (A @ A.mT / 50 + torch.eye(50, device=DEV, dtype=torch.float64)).requires_grad_(True)
    rhs = torch.randn(50, 2, device=DEV, dtype=torch.float64, generator=gen).requires_grad_(True)
    with settings.max_cholesky_size(0):
        sol = DenseLinearOperator(A).solve(rhs)
        iq = DenseLinearOperator(A).inv_quad(rhs)
    assert sol.grad_fn is not None and iq.grad_fn is not None
    iq.backward()
    assert A.grad is not None and rhs.grad is not None

    class Custom(LinearOperator):
        def __init__(self, t):
            super().__init__(t)
            self.t = t

        def _matmul(self, r):
            return _kernels.dense_matmul(self.t, r)

        def _size(self):
            return self.t.shape

        def _transpose_nonbatch(self):
            return self

    A2 = A.detach().clone().requires_grad_(True)
    with settings.max_cholesky_size(0):
        out = Custom(A2).inv_quad(rhs.detach())
    with pytest.raises(NotImplementedError, match="_bilinear_derivative"):
        out.backward()

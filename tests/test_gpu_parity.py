"""GPU parity tests: the CUDA path (through the C ABI, via the reference-shaped Python API) against the golden
fixtures produced by the reference itself and against the numpy oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): index work bit-exact; solves / logdet within 1e-4 relative in fp32 and 1e-10
relative in fp64 *versus the reference's output on identical inputs*.  Every fp64 comparison against the reference or
the oracle is asserted at 1e-10 (measured on B200: <= 6e-11, most at 1e-15; profiles/r2_parity_ledger.md); where a
different bound is used the reason is stated next to it.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import linear_operator_b200 as lo  # noqa: E402
from linear_operator_b200 import settings  # noqa: E402
from linear_operator_b200.operators import (  # noqa: E402
    AddedDiagLinearOperator,
    DenseLinearOperator,
    DiagLinearOperator,
    KroneckerProductLinearOperator,
    LowRankRootLinearOperator,
    ToeplitzLinearOperator,
)
from linear_operator_b200.utils import linear_cg  # noqa: E402
from oracle import krylov_oracle as ko  # noqa: E402

DEV = "cuda:0"
F64_RTOL = 1e-10
F32_RTOL = 1e-4


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a)).to(DEV)


def npy(t):
    return t.detach().cpu().numpy()


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def check(got, want, tol, what=""):
    """Asserts relerr(got, want) < tol and records the MEASURED error next to the bar (conftest.parity_log), so the
    parity table in profiles/ shows how far below the bar each comparison sits."""
    import inspect

    from conftest import parity_log

    if not what:
        ctx = inspect.stack()[1].code_context
        what = ctx[0].strip() if ctx else ""
    err = relerr(got, want)
    parity_log(what, err, tol)
    assert err < tol, f"{what}: relative error {err:.3e} exceeds the bar {tol:.1e}"


class Injected(AddedDiagLinearOperator):
    """AddedDiag with probe vectors handed in through the reference's own hook (_probe_vectors_and_norms)."""

    probes = None

    def _probe_vectors_and_norms(self):
        return self.probes, torch.ones_like(self.probes[..., :1, :])


# ------------------------------------------------------------------------------------------------------------
# linear_cg
# ------------------------------------------------------------------------------------------------------------
def test_cg_vector_fp64(golden):
    g = golden("cg_vec_f64")
    x = linear_cg(cu(g["A"]), cu(g["rhs"]), max_iter=int(g["max_iter"]), tolerance=float(g["tolerance"]))
    assert x.shape == g["x"].shape
    check(npy(x), g["x"], F64_RTOL)


def test_cg_batch_tridiag_fp64(golden):
    g = golden("cg_batch_tridiag_f64")
    x, t = linear_cg(cu(g["A"]), cu(g["rhs"]), n_tridiag=int(g["n_tridiag"]), max_iter=int(g["max_iter"]),
                     max_tridiag_iter=int(g["max_tridiag_iter"]), tolerance=float(g["tolerance"]))
    assert t.shape == g["t_mat"].shape
    check(npy(x), g["x"], F64_RTOL)
    check(npy(t), g["t_mat"], F64_RTOL)


def test_cg_defaults_fp32(golden):
    g = golden("cg_defaults_f32")
    x, t = linear_cg(cu(g["A"]), cu(g["rhs"]), n_tridiag=int(g["n_tridiag"]))
    assert t.shape == g["t_mat"].shape  # same truncation point of the tridiagonal as the reference
    check(npy(x), g["x"], F32_RTOL)
    # late Lanczos coefficients of a converged fp32 run are dominated by round-off in the reference itself
    check(npy(t)[..., :12, :12], g["t_mat"][..., :12, :12], 1e-3)


def test_cg_precond_closure_zero_column_warm_start(golden):
    g = golden("cg_precond_f64")
    minv = cu(g["minv"]).unsqueeze(-1)
    x, t = linear_cg(cu(g["A"]), cu(g["rhs"]), n_tridiag=int(g["n_tridiag"]), max_iter=int(g["max_iter"]),
                     max_tridiag_iter=int(g["max_tridiag_iter"]), tolerance=float(g["tolerance"]),
                     initial_guess=cu(g["x0"]), preconditioner=lambda v: v * minv)
    check(npy(x), g["x"], F64_RTOL)
    check(npy(t), g["t_mat"], F64_RTOL)


def test_cg_identity_truncated_tridiag(golden):
    g = golden("cg_identity_f64")
    x, t = linear_cg(cu(g["A"]), cu(g["rhs"]), n_tridiag=int(g["n_tridiag"]), max_iter=int(g["max_iter"]),
                     max_tridiag_iter=int(g["max_tridiag_iter"]), tolerance=float(g["tolerance"]))
    assert t.shape == g["t_mat"].shape
    check(npy(x), g["x"], 1e-12)
    check(npy(t), g["t_mat"], 1e-12)


def test_cg_python_closure_and_errors():
    torch.manual_seed(0)
    a = torch.randn(40, 40, dtype=torch.float64, device=DEV)
    a = a @ a.mT / 40 + torch.eye(40, dtype=torch.float64, device=DEV)
    b = torch.randn(40, 3, dtype=torch.float64, device=DEV)
    x = linear_cg(lambda v: a @ v, b, max_iter=100, tolerance=1e-10)  # foreign closure: reference route incl. A @ 0
    ref = ko.linear_cg(lambda v: npy(a) @ v, npy(b), max_iter=100, tolerance=1e-10)
    check(npy(x), ref, F64_RTOL)
    with pytest.raises(RuntimeError, match="tridiagonalization larger"):
        linear_cg(a, b, n_tridiag=1, max_iter=3, max_tridiag_iter=5)
    bad = a.clone()
    bad[0, 0] = float("nan")
    with pytest.raises(RuntimeError, match="NaNs encountered"):
        linear_cg(bad, b, max_iter=20)


def test_cg_nonconvergence_warns():
    torch.manual_seed(1)
    a = torch.randn(60, 60, dtype=torch.float64, device=DEV)
    a = a @ a.mT + 1e-3 * torch.eye(60, dtype=torch.float64, device=DEV)
    b = torch.randn(60, 2, dtype=torch.float64, device=DEV)
    with pytest.warns(lo.utils.warnings.NumericalWarning):
        linear_cg(a, b, max_iter=25, tolerance=1e-12)


# ------------------------------------------------------------------------------------------------------------
# pivoted Cholesky + preconditioner
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,rtol", [("pivchol_rbf_f64", F64_RTOL), ("pivchol_rbf_f32", F32_RTOL), ("pivchol_lowrank_f64", F64_RTOL)])
def test_pivoted_cholesky_dense(golden, name, rtol):
    g = golden(name)
    L, perm = lo.pivoted_cholesky(cu(g["A"]), int(g["rank"]), error_tol=float(g["tol"]), return_pivots=True)
    assert perm.dtype == torch.int64
    assert tuple(L.shape) == g["L"].shape  # same (data dependent) rank as the reference
    np.testing.assert_array_equal(npy(perm), g["perm"])  # bit-exact index work
    check(npy(L), g["L"], rtol)


@pytest.mark.parametrize("name", ["precond_const_f64", "precond_varying_f64"])
def test_added_diag_preconditioner(golden, name):
    g = golden(name)
    op = AddedDiagLinearOperator(DenseLinearOperator(cu(g["A"])), DiagLinearOperator(cu(g["d"])))
    with settings.min_preconditioning_size(4), settings.max_preconditioner_size(int(g["rank"])):
        closure, precond_lt, logdet_p = op._preconditioner()
    check(npy(op._piv_chol_self), g["L"], F64_RTOL)
    check(npy(closure(cu(g["v"]))), g["minv_v"], F64_RTOL)
    check(npy(logdet_p), g["logdet_p"], F64_RTOL)
    assert type(precond_lt).__name__ == "PsdSumLinearOperator"


# ------------------------------------------------------------------------------------------------------------
# inv_quad_logdet end to end
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize(
    "name,rtol",
    [("iqld_dense_noprecond_f64", F64_RTOL), ("iqld_dense_precond_f64", F64_RTOL), ("iqld_dense_precond_f32", F32_RTOL),
     ("iqld_dense_noprecond_f32", F32_RTOL)],
)
def test_inv_quad_logdet_dense(golden, name, rtol):
    g = golden(name)
    op = Injected(DenseLinearOperator(cu(g["A"])), DiagLinearOperator(cu(g["d"])))
    op.probes = cu(g["probes"])
    precond = bool(int(g["precond"]))
    with settings.max_cholesky_size(0), settings.min_preconditioning_size(4 if precond else 10**6), \
            settings.max_preconditioner_size(int(g["rank"])):
        iq, ld = op.inv_quad_logdet(cu(g["rhs"]), logdet=True)
        iq_nr, _ = op.inv_quad_logdet(cu(g["rhs"]), logdet=True, reduce_inv_quad=False)
    assert iq.shape == g["inv_quad"].shape and ld.shape == g["logdet"].shape
    check(npy(iq), g["inv_quad"], rtol)
    check(npy(ld), g["logdet"], rtol)
    check(npy(iq_nr), g["inv_quad_noreduce"], rtol)


def test_inv_quad_logdet_rng_probe_stream(golden):
    """Probes drawn by torch.randn in the reference's order: on CUDA the RNG stream differs from the CPU fixture, so
    the check is against the oracle fed with the very base samples the CUDA generator produced."""
    g = golden("iqld_dense_rng_f64")
    A, d, rhs = cu(g["A"]), cu(g["d"]), cu(g["rhs"])
    s, k = g["eps_root"].shape[-1], int(g["rank"])
    op = AddedDiagLinearOperator(DenseLinearOperator(A), DiagLinearOperator(d))
    with settings.max_cholesky_size(0), settings.min_preconditioning_size(4), settings.max_preconditioner_size(k), \
            settings.num_trace_samples(s):
        torch.manual_seed(99)
        iq, ld = op.inv_quad_logdet(rhs, logdet=True)
    torch.manual_seed(99)
    eps_root = torch.randn(2, k, s, dtype=torch.float64, device=DEV)
    eps_diag = torch.randn(s, 2, 60, dtype=torch.float64, device=DEV)
    iq_o, ld_o, _ = ko.dense_added_diag_inv_quad_logdet(g["A"], g["d"], g["rhs"], base_samples=(npy(eps_root), npy(eps_diag)),
                                                       precond_rank=k, min_precond_size=4)
    check(npy(iq), iq_o, F64_RTOL)
    check(npy(ld), ld_o, F64_RTOL)


def test_small_operator_takes_dense_cholesky_path(golden):
    g = golden("iqld_dense_noprecond_f64")
    A, d = g["A"], g["d"]
    op = AddedDiagLinearOperator(DenseLinearOperator(cu(A)), DiagLinearOperator(cu(d)))
    iq, ld = op.inv_quad_logdet(cu(g["rhs"]), logdet=True)  # N=60 <= max_cholesky_size
    full = A + np.eye(60) * d[..., None, :]
    sol = np.linalg.solve(full, g["rhs"])
    check(npy(iq), (sol * g["rhs"]).sum(-2).sum(-1), F64_RTOL)
    check(npy(ld), np.linalg.slogdet(full)[1], F64_RTOL)


def test_torch_function_dispatch(golden):
    g = golden("iqld_dense_noprecond_f64")
    op = AddedDiagLinearOperator(DenseLinearOperator(cu(g["A"])), DiagLinearOperator(cu(g["d"])))
    rhs = cu(g["rhs"])
    dense = g["A"] + np.eye(60) * g["d"][..., None, :]
    check(npy(torch.matmul(op, rhs)), dense @ g["rhs"], 1e-12)
    check(npy(op @ rhs), dense @ g["rhs"], 1e-12)
    with settings.max_cholesky_size(0), settings.cg_tolerance(1e-10), settings.max_cg_iterations(200):
        sol = torch.linalg.solve(op, rhs)
    check(npy(sol), np.linalg.solve(dense, g["rhs"]), 1e-6)  # CG's own eps rule stalls near 1e-6
    check(npy(torch.diagonal(op, dim1=-2, dim2=-1)), np.diagonal(dense, axis1=-1, axis2=-2), 1e-15)
    with pytest.raises(NotImplementedError):
        torch.trace(op)


# ------------------------------------------------------------------------------------------------------------
# SLQ
# ------------------------------------------------------------------------------------------------------------
def test_tridiag_eigh_and_slq(golden):
    g = golden("slq_f64")
    t = cu(g["t_mat"])
    evals, evecs = lo.utils.lanczos.lanczos_tridiag_to_diag(t)
    check(npy(evals), g["evals"], 1e-10)
    check(np.abs(npy(evecs)), np.abs(g["evecs"]), 1e-7)  # signs of eigenvectors are not unique
    ld = lo.utils.StochasticLQ.logdet_from_tridiag(t, int(g["n"]))
    check(npy(ld), g["logdet"], 1e-10)
    (ld2,) = lo.utils.StochasticLQ().to_dense(torch.Size((40, 40)), evals, evecs, [lambda x: x.log()])
    check(npy(ld2), g["logdet"], 1e-10)


# ------------------------------------------------------------------------------------------------------------
# structured operators
# ------------------------------------------------------------------------------------------------------------
def test_kronecker(golden):
    g = golden("kron_f64")
    op = KroneckerProductLinearOperator(cu(g["f0"]), cu(g["f1"]), cu(g["f2"]))
    check(npy(op._matmul(cu(g["x"]))), g["y"], 1e-12)
    check(npy(op._diagonal()), g["diag"], 1e-13)


@pytest.mark.parametrize("name,rtol", [("toeplitz_f64", 1e-12), ("toeplitz_f32", 2e-5)])
def test_toeplitz(golden, name, rtol):
    g = golden(name)
    op = ToeplitzLinearOperator(cu(g["col"]))
    check(npy(op._matmul(cu(g["x"]))), g["y"], rtol)


def test_lowrank_woodbury(golden):
    g = golden("lowrank_f64")
    op = LowRankRootLinearOperator(cu(g["U"])) + DiagLinearOperator(cu(g["d"]))
    assert type(op).__name__ == "LowRankRootAddedDiagLinearOperator"
    check(npy(op.solve(cu(g["rhs"]))), g["solve"], F64_RTOL)
    iq, ld = op.inv_quad_logdet(cu(g["rhs"]), logdet=True)
    check(npy(iq), g["inv_quad"], F64_RTOL)
    check(npy(ld), g["logdet"], F64_RTOL)


def _structured(g, base):
    op = Injected(base, DiagLinearOperator(cu(g["d"])))
    op.probes = cu(g["probes"])
    with settings.max_cholesky_size(0), settings.min_preconditioning_size(4), \
            settings.max_preconditioner_size(int(g["rank"])):
        iq, ld = op.inv_quad_logdet(cu(g["rhs"]), logdet=True)
    check(npy(op._piv_chol_self), g["L"], F64_RTOL)
    check(npy(iq), g["inv_quad"], F64_RTOL)
    check(npy(ld), g["logdet"], F64_RTOL)


def test_inv_quad_logdet_kronecker(golden):
    g = golden("iqld_kron_f64")
    _structured(g, KroneckerProductLinearOperator(cu(g["f0"]), cu(g["f1"]), cu(g["f2"])))


def test_inv_quad_logdet_toeplitz(golden):
    g = golden("iqld_toeplitz_f64")
    _structured(g, ToeplitzLinearOperator(cu(g["col"])))


# ------------------------------------------------------------------------------------------------------------
# larger, oracle-checked and property-checked cases
# ------------------------------------------------------------------------------------------------------------
def _synthetic_dense(B, N, dtype, seed=1234, rank=64):
    gen = torch.Generator(device=DEV).manual_seed(seed)
    W = torch.randn(B, N, rank, dtype=dtype, device=DEV, generator=gen)
    sc = torch.logspace(0, -1.5, rank, dtype=dtype, device=DEV)
    W = W * sc / sc.norm()
    K = W @ W.mT
    d = torch.full((B, N), 0.5, dtype=dtype, device=DEV)
    rhs = torch.randn(B, N, 1, dtype=dtype, device=DEV, generator=gen)
    probes = torch.randn(B, N, 8, dtype=dtype, device=DEV, generator=gen)
    probes = probes / probes.norm(dim=-2, keepdim=True)
    return K, d, rhs, probes


@pytest.mark.parametrize("dtype,rtol", [(torch.float32, F32_RTOL), (torch.float64, F64_RTOL)])
def test_inv_quad_logdet_mid_size_vs_oracle(dtype, rtol):
    """N=700 (ragged against every tile size), batch 3, rank-20 preconditioner: oracle on the same inputs."""
    K, d, rhs, probes = _synthetic_dense(3, 700, dtype)
    op = Injected(DenseLinearOperator(K), DiagLinearOperator(d))
    op.probes = probes
    with settings.max_cholesky_size(0), settings.min_preconditioning_size(100), settings.max_preconditioner_size(20):
        iq, ld = op.inv_quad_logdet(rhs, logdet=True)
    iq_o, ld_o, _ = ko.dense_added_diag_inv_quad_logdet(npy(K), npy(d), npy(rhs), probes=npy(probes), precond_rank=20,
                                                       min_precond_size=100)
    check(npy(iq), iq_o, rtol)
    check(npy(ld), ld_o, rtol)


def test_solve_residual_property_large():
    """Size-independent property at a larger size: A x = b to CG accuracy, and linearity of the solve."""
    K, d, rhs, _ = _synthetic_dense(2, 3000, torch.float32)
    op = AddedDiagLinearOperator(DenseLinearOperator(K), DiagLinearOperator(d))
    with settings.cg_tolerance(1e-4), settings.max_cg_iterations(200), settings.max_preconditioner_size(30):
        x = op.solve(rhs)
        x2 = op.solve(2.5 * rhs)
    res = (op @ x - rhs).norm() / rhs.norm()
    assert res.item() < 1e-3
    check(npy(x2), 2.5 * npy(x), 1e-4)


def test_toeplitz_matmul_batch_chunking_is_exact():
    """BASELINE config 4 needs the FFT product in batch chunks (scratch); chunking must not change a bit."""
    from linear_operator_b200 import _kernels

    gen = torch.Generator(device=DEV).manual_seed(2)
    B, N, C = 7, 300, 5
    col = torch.exp(-0.5 * (torch.arange(N, device=DEV) / 9.0) ** 2).repeat(B, 1) * (1 + torch.arange(B, device=DEV)[:, None])
    X = torch.randn(B, N, C, device=DEV, generator=gen)
    d = 0.5 + torch.rand(B, N, device=DEV, generator=gen)
    ref = _kernels.toeplitz_matmul(col, X, d)
    old = _kernels.TOEPLITZ_SCRATCH_BYTES
    try:
        _kernels.TOEPLITZ_SCRATCH_BYTES = 2 * C * 1024 * 4  # two batch elements per chunk (L = 1024)
        out = _kernels.toeplitz_matmul(col, X, d)
    finally:
        _kernels.TOEPLITZ_SCRATCH_BYTES = old
    assert torch.equal(out, ref)
    dense = torch.stack([torch.stack([col[b].roll(i)[:N] for i in range(N)]) for b in range(B)])  # circulant rows
    idx = (torch.arange(N, device=DEV)[:, None] - torch.arange(N, device=DEV)[None, :]).abs()
    T = col[:, idx]
    want = T.double() @ X.double() + d.double().unsqueeze(-1) * X.double()
    check(npy(out), npy(want), 2e-5)


@pytest.mark.parametrize("dtype,rtol", [(torch.float64, 1e-10), (torch.float32, 1e-4)])
def test_lowrank_shared_root_constant_diag(dtype, rtol):
    """BASELINE config 5's shape of problem: one root U shared by the batch, sigma_b per batch element.  The fast path
    (Gram once, batch as the GEMM row dimension) against a dense solve / logdet of U U^T + sigma_b I."""
    from linear_operator_b200.operators import ConstantDiagLinearOperator

    gen = torch.Generator(device=DEV).manual_seed(8)
    B, N, r = 5, 400, 7
    U = torch.randn(N, r, device=DEV, generator=gen, dtype=dtype) / 3
    sig = 0.5 + torch.rand(B, 1, device=DEV, generator=gen, dtype=dtype)
    rhs = torch.randn(B, N, 1, device=DEV, generator=gen, dtype=dtype)
    op = LowRankRootLinearOperator(U) + ConstantDiagLinearOperator(sig, diag_shape=N)
    assert type(op).__name__ == "LowRankRootAddedDiagLinearOperator" and op._shared_root_constant_diag()
    x = op.solve(rhs)
    iq, ld = op.inv_quad_logdet(rhs, logdet=True)
    dense = (U.double() @ U.double().mT).unsqueeze(0) + sig.double().unsqueeze(-1) * torch.eye(N, device=DEV, dtype=torch.float64)
    xd = torch.linalg.solve(dense, rhs.double())
    check(npy(x), npy(xd), rtol)
    check(npy(iq), npy((rhs.double() * xd).sum((-2, -1))), rtol)
    check(npy(ld), npy(torch.logdet(dense)), rtol)
    # same numbers as the general (batched-root) path
    op2 = LowRankRootLinearOperator(U.expand(B, N, r).contiguous()) + DiagLinearOperator(sig.expand(B, N).contiguous())
    check(npy(op2.solve(rhs)), npy(x), rtol)


def test_inv_quad_logdet_baseline_operator_size_vs_oracle():
    """BASELINE config 2's operator exactly (N = 5000, 256-direction decaying spectrum + 0.5 I, fp32, 32 probes + 1 rhs =
    33 columns, rank-100 pivoted-Cholesky preconditioner, default settings = 21 CG iterations), batch 2: the whole CUDA
    path (streaming tcgen05 matmul, fused CG updates, preconditioner, SLQ) against the numpy oracle on identical inputs
    and probes.  This is where the tensor core's truncating accumulate would show if it mattered: bar 1e-4 (north star)."""
    gen = torch.Generator(device=DEV).manual_seed(1234)
    B, N, S = 2, 5000, 32
    W = torch.randn(B, N, 256, device=DEV, generator=gen)
    sc = torch.logspace(0, -1.5, 256, device=DEV)
    W = W * sc / sc.norm()
    K = W @ W.mT
    d = torch.full((B, N), 0.5, device=DEV)
    rhs = torch.randn(B, N, 1, device=DEV, generator=gen)
    probes = torch.randn(B, N, S, device=DEV, generator=gen)
    probes = probes / probes.norm(dim=-2, keepdim=True)
    op = Injected(DenseLinearOperator(K), DiagLinearOperator(d))
    op.probes = probes
    with settings.num_trace_samples(S), settings.max_preconditioner_size(100):
        iq, ld = op.inv_quad_logdet(rhs, logdet=True)
    iq_o, ld_o, _ = ko.dense_added_diag_inv_quad_logdet(npy(K), npy(d), npy(rhs), probes=npy(probes), precond_rank=100)
    e_iq = np.abs(npy(iq) - iq_o) / np.abs(iq_o)
    e_ld = np.abs(npy(ld) - ld_o) / np.abs(ld_o)
    print(f"N=5000 vs oracle: inv_quad rel err {e_iq.max():.2e}, logdet rel err {e_ld.max():.2e}")
    assert e_iq.max() < F32_RTOL and e_ld.max() < F32_RTOL


@pytest.mark.parametrize("name,rt", [("entry_dense_f64", F64_RTOL / 10), ("entry_dense_f32", F32_RTOL / 10)])
def test_solve_and_inv_quad_entry_points(golden, name, rt):
    """SURVEY 8a row 17 against the reference's own outputs: op.solve (with and without left tensor, vector rhs on an
    un-batched operator), torch.linalg.solve dispatch, op.inv_quad (reduced / not), default (loose) tolerance."""
    g = golden(name)
    A, d, rhs = cu(g["A"]), cu(g["d"]), cu(g["rhs"])
    op = AddedDiagLinearOperator(DenseLinearOperator(A), DiagLinearOperator(d))
    rank, tol = int(g["rank"]), float(g["tol"])
    with settings.max_cholesky_size(0), settings.min_preconditioning_size(4), settings.max_preconditioner_size(rank), \
            settings.cg_tolerance(tol), settings.max_cg_iterations(300):
        sol = op.solve(rhs)
        sol_l = op.solve(rhs, cu(g["lhs"]))
        iq = op.inv_quad(rhs)
        iq_nr = op.inv_quad(rhs, reduce_inv_quad=False)
        sol_t = torch.linalg.solve(op, rhs)
        op1 = AddedDiagLinearOperator(DenseLinearOperator(A[0]), DiagLinearOperator(d[0]))
        sol_v = op1.solve(cu(g["vec"]))
    with settings.max_cholesky_size(0), settings.min_preconditioning_size(4), settings.max_preconditioner_size(rank):
        sol_def = op.solve(rhs)
    assert sol.shape == g["solve"].shape and sol_l.shape == g["solve_left"].shape and sol_v.shape == g["solve_vec"].shape
    assert iq.shape == g["inv_quad"].shape and iq_nr.shape == g["inv_quad_noreduce"].shape
    check(npy(sol), g["solve"], rt)
    check(npy(sol_t), g["solve_torch"], rt)
    check(npy(sol_l), g["solve_left"], 10 * rt)
    check(npy(iq), g["inv_quad"], 10 * rt)
    check(npy(iq_nr), g["inv_quad_noreduce"], 10 * rt)
    check(npy(sol_v), g["solve_vec"], 10 * rt)
    check(npy(sol_def), g["solve_default"], 10 * rt)


def test_add_jitter_constant_diag_entry_points(golden):
    """DenseLinearOperator.add_jitter -> AddedDiag(Dense, ConstantDiag) (stride-0 diagonal reaches the kernels):
    inv_quad_logdet, logdet alone and solve against the reference's outputs."""
    g = golden("entry_jitter_f64")
    base = DenseLinearOperator(cu(g["A"])).add_jitter(float(g["jitter"]))
    assert type(base).__name__ == "AddedDiagLinearOperator" and type(base._diag_tensor).__name__ == "ConstantDiagLinearOperator"
    op = Injected(base._linear_op, base._diag_tensor)
    op.probes = cu(g["probes"])
    rhs = cu(g["rhs"])
    with settings.max_cholesky_size(0), settings.min_preconditioning_size(4), settings.max_preconditioner_size(int(g["rank"])):
        iq, ld = op.inv_quad_logdet(rhs, logdet=True)
        ld_only = op.logdet()
        with settings.cg_tolerance(1e-8), settings.max_cg_iterations(300):
            sol = base.solve(rhs)
    check(npy(iq), g["inv_quad"], F64_RTOL)
    check(npy(ld), g["logdet"], F64_RTOL)
    check(npy(ld_only), g["logdet_only"], F64_RTOL)
    check(npy(sol), g["solve"], F64_RTOL)

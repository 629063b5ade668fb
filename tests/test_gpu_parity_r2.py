"""GPU parity tests, round 2: the BASELINE shapes of configs 1 / 3 / 4 / 5, fp32 Kronecker, bit-exact pivots of
structured and generic operators, generic-operator pivoted Cholesky, N = 4096 Toeplitz.

Bars (BASELINE.json north_star): index work bit-exact; 1e-10 relative in fp64 and 1e-4 relative in fp32 against the
reference's outputs (fixtures of tests/golden/make_golden_round2.py) or against the oracle on identical inputs.
Every comparison records its measured error in the parity ledger (conftest.parity_log).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import linear_operator_b200 as lo  # noqa: E402
from linear_operator_b200 import settings  # noqa: E402
from linear_operator_b200.operators import (  # noqa: E402
    AddedDiagLinearOperator,
    ConstantDiagLinearOperator,
    DenseLinearOperator,
    DiagLinearOperator,
    KroneckerProductLinearOperator,
    LowRankRootLinearOperator,
    RootLinearOperator,
    ToeplitzLinearOperator,
)
from oracle import krylov_oracle as ko  # noqa: E402
from test_gpu_parity import DEV, F32_RTOL, F64_RTOL, Injected, check, cu, npy  # noqa: E402


# ------------------------------------------------------------------------------------------------------------
# BASELINE configs[0]: Dense + AddedDiag, N = 512, batch 1, fp64, 16 probes
# ------------------------------------------------------------------------------------------------------------
def test_cfg1_baseline_shape_vs_reference(golden):
    g = golden("cfg1_dense_f64")
    W = cu(g["W"])
    op = Injected(DenseLinearOperator(W @ W.mT), DiagLinearOperator(cu(g["d"])))
    op.probes = cu(g["probes"])
    with settings.max_cholesky_size(0), settings.num_trace_samples(16):
        iq, ld = op.inv_quad_logdet(cu(g["rhs"]), logdet=True)
        sol = op.solve(cu(g["rhs"]))
    assert iq.shape == g["inv_quad"].shape and ld.shape == g["logdet"].shape
    check(npy(iq), g["inv_quad"], F64_RTOL)
    check(npy(ld), g["logdet"], F64_RTOL)
    check(npy(sol), g["solve"], F64_RTOL)


# ------------------------------------------------------------------------------------------------------------
# Kronecker in fp32 (reference fixtures) and at BASELINE configs[2]'s shape (oracle)
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["kron_f32", "kron_mid_f32"])
def test_kronecker_fp32_vs_reference(golden, name):
    g = golden(name)
    op = KroneckerProductLinearOperator(cu(g["f0"]), cu(g["f1"]), cu(g["f2"]))
    y = op._matmul(cu(g["x"]))
    assert y.dtype == torch.float32
    check(npy(y), g["y"], F32_RTOL)
    if "diag" in g:
        check(npy(op._diagonal()), g["diag"], 1e-6)


def test_inv_quad_logdet_kronecker_fp32_vs_reference(golden):
    g = golden("iqld_kron_f32")
    kron = KroneckerProductLinearOperator(cu(g["f0"]), cu(g["f1"]), cu(g["f2"]))
    op = Injected(kron, DiagLinearOperator(cu(g["d"])))
    op.probes = cu(g["probes"])
    with settings.max_cholesky_size(0), settings.min_preconditioning_size(4), settings.max_preconditioner_size(5):
        iq, ld = op.inv_quad_logdet(cu(g["rhs"]), logdet=True)
        L, perm = kron.pivoted_cholesky(rank=5, return_pivots=True)
    np.testing.assert_array_equal(npy(perm), g["perm"])
    check(npy(L), g["L"], F32_RTOL)
    check(npy(iq), g["inv_quad"], F32_RTOL)
    check(npy(ld), g["logdet"], F32_RTOL)


def test_cfg3_kronecker_baseline_shape_vs_oracle():
    """BASELINE configs[2]: three 100 x 100 factors (N = 10^6) + 0.5 I, fp32, 32 injected probes, default settings
    (rank-15 preconditioner, 21 CG iterations); one batch element, against the oracle on identical inputs."""
    gen = torch.Generator(device=DEV).manual_seed(77)
    fs = []
    for _ in range(3):
        G = torch.randn(100, 100, device=DEV, generator=gen)
        fs.append(G @ G.mT / 100 + 0.1 * torch.eye(100, device=DEV))
    N, S = 10**6, 32
    d = torch.full((N,), 0.5, device=DEV)
    rhs = torch.randn(N, 1, device=DEV, generator=gen)
    probes = torch.randn(N, S, device=DEV, generator=gen)
    probes = probes / probes.norm(dim=-2, keepdim=True)
    kron = KroneckerProductLinearOperator(*fs)
    op = Injected(kron, DiagLinearOperator(d))
    op.probes = probes
    with settings.num_trace_samples(S):
        iq, ld = op.inv_quad_logdet(rhs, logdet=True)
        L, perm = kron.pivoted_cholesky(rank=15, return_pivots=True)
    fsn = [npy(f) for f in fs]
    Lo, permo = ko.pivoted_cholesky(ko.kron_diag(fsn), lambda pi: ko.kron_rows(fsn, pi), 15)
    np.testing.assert_array_equal(npy(perm), permo)  # bit-exact pivots at N = 10^6
    check(npy(L), Lo, F32_RTOL)
    # At N = 10^6 the fp32 oracle is itself noisy (numpy reduces the 10^6-row column dots sequentially in fp32: measured
    # 3.5e-4 away from this path), so the checker is the oracle run in fp64 on the SAME fp32 inputs: the CUDA path must
    # be within the fp32 bar of that, which is the stronger statement.
    f64 = [f.astype(np.float64) for f in fsn]
    X = torch.cat([probes, rhs], -1)
    check(npy(kron._matmul(X)), ko.kron_matmul(f64, npy(X).astype(np.float64)), 1e-5)
    dn = npy(d).astype(np.float64)
    closure, logdet_p, _ = ko.added_diag_preconditioner(Lo.astype(np.float64), dn)
    iq_o, ld_o, _ = ko.inv_quad_logdet(lambda v: ko.kron_matmul(f64, v) + dn[..., None] * v, N,
                                       npy(rhs).astype(np.float64), npy(probes).astype(np.float64), closure, logdet_p)
    check(npy(iq), iq_o, F32_RTOL)
    check(npy(ld), ld_o, F32_RTOL)


# ------------------------------------------------------------------------------------------------------------
# bit-exact pivots of structured and generic operators
# ------------------------------------------------------------------------------------------------------------
def test_pivots_kronecker_toeplitz_bit_exact(golden):
    g = golden("pivchol_kron_f64")
    L, perm = KroneckerProductLinearOperator(cu(g["f0"]), cu(g["f1"]), cu(g["f2"])).pivoted_cholesky(
        rank=int(g["rank"]), return_pivots=True)
    np.testing.assert_array_equal(npy(perm), g["perm"])
    check(npy(L), g["L"], F64_RTOL)
    g = golden("pivchol_toeplitz_f64")
    L, perm = ToeplitzLinearOperator(cu(g["col"])).pivoted_cholesky(rank=int(g["rank"]), return_pivots=True)
    np.testing.assert_array_equal(npy(perm), g["perm"])
    check(npy(L), g["L"], F64_RTOL)
    g = golden("iqld_kron_f64")
    L, perm = KroneckerProductLinearOperator(cu(g["f0"]), cu(g["f1"]), cu(g["f2"])).pivoted_cholesky(
        rank=int(g["rank"]), return_pivots=True)
    fs = [g["f0"], g["f1"], g["f2"]]
    _, permo = ko.pivoted_cholesky(ko.kron_diag(fs), lambda pi: ko.kron_rows(fs, pi), int(g["rank"]))
    np.testing.assert_array_equal(npy(perm), permo)
    g = golden("iqld_toeplitz_f64")
    col = g["col"]
    L, perm = ToeplitzLinearOperator(cu(col)).pivoted_cholesky(rank=int(g["rank"]), return_pivots=True)
    _, permo = ko.pivoted_cholesky(np.full(80, col[0]), lambda pi: ko.sym_toeplitz_rows(col, pi), int(g["rank"]))
    np.testing.assert_array_equal(npy(perm), permo)


def test_generic_pivoted_cholesky_root_and_sum(golden):
    """Operators without a device row functor take the reference's generic route (rows through _get_indices)."""
    g = golden("pivchol_root_f64")
    L, perm = RootLinearOperator(cu(g["U"])).pivoted_cholesky(rank=int(g["rank"]), return_pivots=True)
    assert tuple(L.shape) == g["L"].shape
    np.testing.assert_array_equal(npy(perm), g["perm"])
    check(npy(L), g["L"], F64_RTOL)
    g = golden("pivchol_sum_f64")
    op = DenseLinearOperator(cu(g["A"])) + RootLinearOperator(cu(g["U"]))
    assert type(op).__name__ == "SumLinearOperator"
    L, perm = op.pivoted_cholesky(rank=int(g["rank"]), return_pivots=True)
    np.testing.assert_array_equal(npy(perm), g["perm"])
    check(npy(L), g["L"], F64_RTOL)


def test_generic_pivoted_cholesky_early_stop_and_user_operator():
    """A user-defined operator (only _matmul/_size/_transpose_nonbatch/_get_indices/_diagonal) of rank 3: the loop stops
    after 3 steps exactly like the dense kernel on the materialised matrix."""
    from linear_operator_b200.operators import LinearOperator

    class Outer(LinearOperator):
        def __init__(self, w):
            super().__init__(w)
            self.w = w

        def _matmul(self, rhs):
            return self.w @ (self.w.mT @ rhs)

        def _size(self):
            return torch.Size((*self.w.shape[:-2], self.w.shape[-2], self.w.shape[-2]))

        def _transpose_nonbatch(self):
            return self

        def _diagonal(self):
            return (self.w * self.w).sum(-1)

        def _get_indices(self, row_index, col_index, *batch_indices):
            return (self.w[(*batch_indices, row_index)] * self.w[(*batch_indices, col_index)]).sum(-1)

    gen = torch.Generator(device=DEV).manual_seed(5)
    w = torch.randn(3, 90, 3, dtype=torch.float64, device=DEV, generator=gen)
    L, perm = Outer(w).pivoted_cholesky(rank=20, error_tol=1e-9, return_pivots=True)
    Ld, permd = DenseLinearOperator(w @ w.mT).pivoted_cholesky(rank=20, error_tol=1e-9, return_pivots=True)
    assert L.shape == Ld.shape and L.shape[-1] <= 4
    assert torch.equal(perm, permd)
    check(npy(L), npy(Ld), 1e-9)


def test_added_diag_root_preconditioned_path(golden):
    """AddedDiag(Root(U), Diag) through pivoted Cholesky + preconditioned mBCG."""
    g = golden("iqld_root_f64")
    op = Injected(RootLinearOperator(cu(g["U"])), DiagLinearOperator(cu(g["d"])))
    op.probes = cu(g["probes"])
    with settings.max_cholesky_size(0), settings.min_preconditioning_size(4), \
            settings.max_preconditioner_size(int(g["rank"])):
        iq, ld = op.inv_quad_logdet(cu(g["rhs"]), logdet=True)
    check(npy(op._piv_chol_self), g["L"], F64_RTOL)
    check(npy(iq), g["inv_quad"], F64_RTOL)
    check(npy(ld), g["logdet"], F64_RTOL)


def test_added_diag_root_large_runs():
    """N >= 2000 (the default min_preconditioning_size): the case the round-1 verdict flagged as crashing."""
    gen = torch.Generator(device=DEV).manual_seed(6)
    U = torch.randn(2, 2500, 40, device=DEV, generator=gen) / 6
    d = torch.full((2, 2500), 0.5, device=DEV)
    rhs = torch.randn(2, 2500, 1, device=DEV, generator=gen)
    op = AddedDiagLinearOperator(RootLinearOperator(U), DiagLinearOperator(d))
    with settings.max_preconditioner_size(10), settings.cg_tolerance(1e-5), settings.max_cg_iterations(200):
        x = op.solve(rhs)
    dense = U.double() @ U.double().mT + torch.diag_embed(d.double())
    check(npy(x), npy(torch.linalg.solve(dense, rhs.double())), F32_RTOL)


# ------------------------------------------------------------------------------------------------------------
# Toeplitz: N = 4096 reference fixtures, and BASELINE configs[3]'s N = 2^20 against the oracle
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,rtol", [("toeplitz_big_f64", F64_RTOL), ("toeplitz_big_f32", F32_RTOL)])
def test_toeplitz_n4096_vs_reference(golden, name, rtol):
    g = golden(name)
    op = ToeplitzLinearOperator(cu(g["col"]))
    check(npy(op._matmul(cu(g["x"]))), g["y"], rtol)


def test_toeplitz_subbatch_column_broadcast():
    """Column batch (3,) against a right-hand side batch (2, 3): the cached spectrum is broadcast, not mis-sliced."""
    gen = torch.Generator(device=DEV).manual_seed(3)
    N = 50
    col = torch.exp(-0.5 * (torch.arange(N, device=DEV, dtype=torch.float64) / 5.0) ** 2).repeat(3, 1)
    col = col * torch.tensor([[1.0], [1.5], [2.0]], device=DEV, dtype=torch.float64)
    X = torch.randn(2, 3, N, 4, device=DEV, dtype=torch.float64, generator=gen)
    y = ToeplitzLinearOperator(col)._matmul(X)
    want = ko.sym_toeplitz_matmul(npy(col), npy(X))
    check(npy(y), want, 1e-12)


def test_toeplitz_matmul_column_pairs_of_very_different_scale():
    """toeplitz_matmul packs two real columns into one complex FFT; each column is scaled by an exact power of two
    first, so a column 10^12 times smaller than its partner keeps its own relative accuracy (per-column check against
    the fp64 oracle), also with an odd column count, an even one and a single column."""
    gen = torch.Generator(device=DEV).manual_seed(17)
    N = 1500
    col = torch.exp(-0.5 * (torch.arange(N, device=DEV) / 7.0) ** 2)
    for C in (5, 4, 1):
        scales = torch.tensor([1e6, 1e-6, 1.0, 1e-3, 1e3][:C], device=DEV)
        X = torch.randn(N, C, device=DEV, generator=gen) * scales
        if C == 5:
            X[:, 2] = 0.0  # an all-zero column beside a non-zero partner
        y = ToeplitzLinearOperator(col)._matmul(X)
        want = ko.sym_toeplitz_matmul(npy(col).astype(np.float64), npy(X).astype(np.float64))
        for c in range(C):
            ref = np.abs(want[:, c]).max()
            err = np.abs(npy(y)[:, c] - want[:, c]).max()
            assert err <= 2e-5 * ref, f"column {c} (scale {scales[c].item():g}): {err / max(ref, 1e-300):.2e}"  # zero column: exactly 0


def test_structured_products_fused_dot_partials():
    """Kronecker chain and Toeplitz column-pair product hand linear_cg the <X, Y> partial sums of their last pass
    (linear_cg.py:250-251): they must add up to the dots of the product they come with, ragged row blocks included."""
    from linear_operator_b200 import _kernels

    gen = torch.Generator(device=DEV).manual_seed(23)
    fs = []
    for n in (8, 12, 20):
        G = torch.randn(2, n, n, device=DEV, generator=gen)
        fs.append(G @ G.mT / n + 0.1 * torch.eye(n, device=DEV))
    N, C = 8 * 12 * 20, 33
    X = torch.randn(2, N, C, device=DEV, generator=gen)
    d = 0.5 + torch.rand(2, N, device=DEV, generator=gen)
    Y, dots, n_parts = _kernels.kron_matmul(fs, X, d=d, want_dots=True)
    assert dots is not None and dots.shape == (2, n_parts, C)
    assert torch.equal(Y, _kernels.kron_matmul(fs, X, d=d))
    check(npy(dots.sum(1)), npy((X.double() * Y.double()).sum(-2)), 1e-6)
    col = torch.exp(-0.5 * (torch.arange(1000, device=DEV) / 9.0) ** 2).repeat(2, 1)
    Xt = torch.randn(2, 1000, C, device=DEV, generator=gen)
    for dd in (None, 0.5 + torch.rand(2, 1000, device=DEV, generator=gen)):
        Yt, dots, n_parts = _kernels.toeplitz_matmul(col, Xt, dd, want_dots=True)
        assert dots.shape == (2, n_parts, C) and torch.equal(Yt, _kernels.toeplitz_matmul(col, Xt, dd))
        check(npy(dots.sum(1)), npy((Xt.double() * Yt.double()).sum(-2)), 1e-6)
    Xd = Xt.double()
    Yd, dots, _ = _kernels.toeplitz_matmul(col.double(), Xd, None, want_dots=True)
    check(npy(dots.sum(1)), npy((Xd * Yd).sum(-2)), 1e-13)


def test_toeplitz_matmul_wide_column_blocks():
    """Column counts beyond the 33 of the solver: 100 columns take 64-row tiles in pack / unpack, 300 columns (more than
    one shared-memory tile holds) are split into column blocks on the host."""
    gen = torch.Generator(device=DEV).manual_seed(19)
    N = 700
    col = torch.exp(-0.5 * (torch.arange(N, device=DEV, dtype=torch.float64) / 6.0) ** 2)
    for C in (100, 300):
        X = torch.randn(N, C, device=DEV, dtype=torch.float64, generator=gen)
        y = ToeplitzLinearOperator(col)._matmul(X)
        check(npy(y), ko.sym_toeplitz_matmul(npy(col), npy(X)), 1e-12)


@pytest.mark.parametrize("N,C", [(1001, 4), (1003, 8), (999, 12), (1000, 33), (1002, 2), (1001, 3)])
def test_linear_cg_float4_vector_kernels_ragged_shapes(N, C):
    """fp32 mBCG through the float4 forms of the per-iteration vector kernels (N * C % 4 == 0): row counts that leave a
    ragged last group of four rows, column counts around the lane width, against the oracle's restatement of the
    reference on identical inputs (same iteration count, tridiagonals included); the last shape (N * C odd) takes the
    scalar kernels."""
    from linear_operator_b200.utils import linear_cg

    gen = torch.Generator(device=DEV).manual_seed(100 + N + C)
    W = torch.randn(2, N, N, device=DEV, generator=gen)  # full-rank spectrum in [0.5, 4.5]: no early breakdown of the
    A = W @ W.mT / N + 0.5 * torch.eye(N, device=DEV)    # Lanczos coefficients, which would amplify fp32 rounding
    rhs = torch.randn(2, N, C, device=DEV, generator=gen)
    nt = min(C, 2)
    x, t = linear_cg(A, rhs, n_tridiag=nt, max_iter=12, max_tridiag_iter=8, tolerance=1e-9)
    xo, to = ko.linear_cg(npy(A), npy(rhs), n_tridiag=nt, max_iter=12, max_tridiag_iter=8, tolerance=1e-9)
    check(npy(x), xo, F32_RTOL)
    check(npy(t), to, F32_RTOL)


def test_cfg4_toeplitz_baseline_shape_vs_oracle():
    """BASELINE configs[3]: toeplitz_matmul at N = 2^20 with the full 33-column block, one batch element, fp32, against
    the oracle's length-(2N-1) complex-FFT restatement (utils/toeplitz.py:131-149)."""
    gen = torch.Generator(device=DEV).manual_seed(9)
    N, C = 2**20, 33
    col = torch.exp(-0.5 * (torch.arange(N, device=DEV) / 50.0) ** 2)
    X = torch.randn(N, C, device=DEV, generator=gen)
    d = torch.full((N,), 0.5, device=DEV)
    op = ToeplitzLinearOperator(col)
    y = op._matmul(X)
    want = ko.sym_toeplitz_matmul(npy(col), npy(X))
    check(npy(y), want, F32_RTOL)
    yd = AddedDiagLinearOperator(op, DiagLinearOperator(d))._matmul(X)
    check(npy(yd), want + 0.5 * npy(X), F32_RTOL)


# ------------------------------------------------------------------------------------------------------------
# BASELINE configs[4]: shared low-rank root at N = 10^6, r = 256
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype,rtol", [(torch.float64, F64_RTOL), (torch.float32, F32_RTOL)])
def test_cfg5_lowrank_shared_root_n1e6_vs_oracle(dtype, rtol):
    gen = torch.Generator(device=DEV).manual_seed(12)
    B, N, r = 4, 10**6, 256
    U = torch.randn(N, r, device=DEV, generator=gen, dtype=dtype) / 16
    sig = (0.5 * (1 + torch.arange(B, device=DEV, dtype=dtype) / B)).reshape(B, 1)
    rhs = torch.randn(B, N, 1, device=DEV, generator=gen, dtype=dtype)
    op = LowRankRootLinearOperator(U) + ConstantDiagLinearOperator(sig, diag_shape=N)
    assert type(op).__name__ == "LowRankRootAddedDiagLinearOperator" and op._shared_root_constant_diag()
    x = op.solve(rhs)
    iq, ld = op.inv_quad_logdet(rhs, logdet=True)
    Un = npy(U).astype(np.float64)
    dn = np.broadcast_to(npy(sig).astype(np.float64), (B, N))
    rn = npy(rhs).astype(np.float64)
    xo = np.stack([ko.lowrank_solve(Un, dn[b], rn[b]) for b in range(B)])
    ldo = np.stack([ko.lowrank_logdet(Un, dn[b]) for b in range(B)])
    check(npy(x), xo, rtol)
    check(npy(iq), (rn * xo).sum((-2, -1)), rtol)
    check(npy(ld), ldo, rtol)


# ------------------------------------------------------------------------------------------------------------
# CUDA-graph replay of small dense solves (settings.cuda_graphs)
# ------------------------------------------------------------------------------------------------------------
def test_cuda_graph_replay_is_bit_identical_to_eager_launches():
    """BASELINE configs[0]-sized problems are launch-bound; linear_cg replays the fixed-length part of the solve as one
    CUDA graph on static buffers.  Same kernels, same order: the results must not differ by a bit, also when the graph
    is replayed on new data and when the stop rule only fires after the captured part."""
    import sys

    cgm = sys.modules["linear_operator_b200.utils.linear_cg"]  # (the package re-exports the function under this name)

    gen = torch.Generator(device=DEV).manual_seed(17)
    n, s = 512, 16
    outs = {}
    for flag in (False, True, True):  # eager, capture + replay, replay on new inputs below
        res = []
        for trial in range(2):
            g2 = torch.Generator(device=DEV).manual_seed(100 + trial)
            W = torch.randn(n, 64, device=DEV, dtype=torch.float64, generator=g2) / 8
            K = W @ W.mT
            d = torch.full((n,), 0.5 + 0.1 * trial, device=DEV, dtype=torch.float64)
            rhs = torch.randn(n, 1, device=DEV, dtype=torch.float64, generator=g2)
            probes = torch.randn(n, s, device=DEV, dtype=torch.float64, generator=g2)
            probes = probes / probes.norm(dim=-2, keepdim=True)
            op = Injected(DenseLinearOperator(K), DiagLinearOperator(d))
            op.probes = probes
            with settings.cuda_graphs(flag), settings.max_cholesky_size(0), settings.num_trace_samples(s):
                iq, ld = op.inv_quad_logdet(rhs, logdet=True)
                with settings.cg_tolerance(1e-9), settings.max_cg_iterations(300):
                    sol = op.solve(rhs)  # the stop rule fires long after the captured iterations
            res.append((iq.clone(), ld.clone(), sol.clone()))
        outs.setdefault(flag, []).append(res)
    assert len(cgm._GRAPHS) >= 1  # the graph path really ran
    eager = outs[False][0]
    for replayed in outs[True]:
        for (a, b, c), (a2, b2, c2) in zip(eager, replayed):
            assert torch.equal(a, a2) and torch.equal(b, b2) and torch.equal(c, c2)
    del gen


# ------------------------------------------------------------------------------------------------------------
# Kronecker + constant diagonal through the Kronecker eigenbasis
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,rtol", [("kron_added_diag_f64", F64_RTOL), ("kron_added_diag_f32", F32_RTOL)])
def test_kronecker_added_diag_eigen_path_vs_reference(golden, name, rtol):
    from linear_operator_b200.operators import KroneckerProductAddedDiagLinearOperator

    g = golden(name)
    op = KroneckerProductLinearOperator(cu(g["f0"]), cu(g["f1"]), cu(g["f2"])).add_jitter(float(g["jitter"]))
    assert type(op) is KroneckerProductAddedDiagLinearOperator and op._preconditioner() == (None, None, None)
    rhs = cu(g["rhs"])
    with settings.max_cholesky_size(0):
        sol = op.solve(rhs)
        iq, ld = op.inv_quad_logdet(rhs, logdet=True)
        ld_only = torch.logdet(op)
    assert sol.dtype == rhs.dtype and ld.dtype == rhs.dtype
    check(npy(sol), g["solve"], rtol)
    check(npy(iq), g["inv_quad"], rtol)
    check(npy(ld), g["logdet"], rtol)
    check(npy(ld_only), g["logdet_only"], rtol)


def test_kronecker_added_diag_eigen_path_mid_size_vs_dense():
    """20 x 20 x 20 (N = 8000, the tensor-core chain's alignment) in fp32, against a dense fp64 solve."""
    gen = torch.Generator(device=DEV).manual_seed(44)
    fs = []
    for _ in range(3):
        G = torch.randn(20, 20, device=DEV, generator=gen)
        fs.append(G @ G.mT / 20 + 0.1 * torch.eye(20, device=DEV))
    op = KroneckerProductLinearOperator(*fs).add_jitter(0.3)
    rhs = torch.randn(8000, 4, device=DEV, generator=gen)
    sol = op.solve(rhs)
    dense = torch.kron(torch.kron(fs[0].double(), fs[1].double()), fs[2].double()) + 0.3 * torch.eye(8000, device=DEV, dtype=torch.float64)
    check(npy(sol), npy(torch.linalg.solve(dense, rhs.double())), F32_RTOL)
    check(npy(op.logdet()), npy(torch.logdet(dense)), F32_RTOL)

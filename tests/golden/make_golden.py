#!/usr/bin/env python3
"""Generates tests/golden/*.npz by importing and RUNNING THE REFERENCE (read-only at /root/reference) on small
seeded inputs.  Run in the build container only (the GPU box has no /root/reference):

    cd /tmp && python /root/repo/tests/golden/make_golden.py

The fixtures are the parity pin for oracle/krylov_oracle.py (tests/test_oracle_golden.py) and the expected
values of the GPU parity tests.  Inputs are stored next to the outputs so nothing has to be regenerated.
"""
import os
import sys
import warnings

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
import linear_operator as lo  # noqa: E402
from linear_operator import settings  # noqa: E402
from linear_operator.operators import (  # noqa: E402
    AddedDiagLinearOperator,
    DenseLinearOperator,
    DiagLinearOperator,
    KroneckerProductLinearOperator,
    LowRankRootLinearOperator,
    ToeplitzLinearOperator,
)
from linear_operator.utils.lanczos import lanczos_tridiag, lanczos_tridiag_to_diag  # noqa: E402
from linear_operator.utils.linear_cg import linear_cg  # noqa: E402
from linear_operator.utils.stochastic_lq import StochasticLQ  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
warnings.simplefilter("ignore")


def npy(t):
    return t.detach().cpu().numpy()


def save(name, **arrs):
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **{k: np.asarray(v) for k, v in arrs.items()})
    print("wrote", name, {k: np.asarray(v).shape for k, v in arrs.items()})


def rbf(n, dtype, batch=(), ls=0.3, gen=None):
    x = torch.rand(*batch, n, 2, dtype=dtype, generator=gen)
    d2 = (x.unsqueeze(-2) - x.unsqueeze(-3)).pow(2).sum(-1)
    return torch.exp(-0.5 * d2 / ls**2)


def wishart(n, dtype, batch=(), rank=None, gen=None, jitter=0.0):
    rank = rank or n
    w = torch.randn(*batch, n, rank, dtype=dtype, generator=gen)
    return w @ w.mT / rank + jitter * torch.eye(n, dtype=dtype)


def cg_cases():
    g = torch.Generator().manual_seed(11)
    # (a) vector rhs, fp64, run to convergence (test/utils/test_linear_cg.py:27-44 style)
    n = 100
    a = wishart(n, torch.float64, gen=g, jitter=1.0)
    b = torch.randn(n, dtype=torch.float64, generator=g)
    x = linear_cg(a.matmul, b, max_iter=200, tolerance=1e-8)
    save("cg_vec_f64", A=npy(a), rhs=npy(b), x=npy(x), max_iter=200, tolerance=1e-8)

    # (b) batch + tridiag, fp64
    a = wishart(30, torch.float64, batch=(5,), gen=g, jitter=0.5)
    b = torch.randn(5, 30, 6, dtype=torch.float64, generator=g)
    x, t = linear_cg(a.matmul, b, n_tridiag=4, max_iter=30, max_tridiag_iter=10, tolerance=1e-6)
    save("cg_batch_tridiag_f64", A=npy(a), rhs=npy(b), x=npy(x), t_mat=npy(t), n_tridiag=4, max_iter=30,
         max_tridiag_iter=10, tolerance=1e-6)

    # (c) defaults (tol=1, 20 Lanczos iterations => exactly 21 iterations), fp32, two batch dims
    a = wishart(64, torch.float32, batch=(2, 3), gen=g, jitter=0.5)
    b = torch.randn(2, 3, 64, 5, dtype=torch.float32, generator=g)
    x, t = linear_cg(a.matmul, b, n_tridiag=3)
    save("cg_defaults_f32", A=npy(a), rhs=npy(b), x=npy(x), t_mat=npy(t), n_tridiag=3)

    # (d) preconditioned (Jacobi closure), zero rhs column, warm start, fp64
    a = wishart(48, torch.float64, batch=(2,), gen=g, jitter=0.2) * torch.linspace(1, 5, 48, dtype=torch.float64)
    a = 0.5 * (a + a.mT) + torch.diag_embed(torch.linspace(0.5, 3, 48, dtype=torch.float64))
    minv = 1.0 / a.diagonal(dim1=-1, dim2=-2)
    b = torch.randn(2, 48, 4, dtype=torch.float64, generator=g)
    b[..., 2] = 0.0
    x0 = 0.1 * torch.randn(2, 48, 4, dtype=torch.float64, generator=g)
    x, t = linear_cg(a.matmul, b, n_tridiag=2, max_iter=40, max_tridiag_iter=12, tolerance=1e-7,
                     initial_guess=x0, preconditioner=lambda v: v * minv.unsqueeze(-1))
    save("cg_precond_f64", A=npy(a), rhs=npy(b), x0=npy(x0), minv=npy(minv), x=npy(x), t_mat=npy(t), n_tridiag=2,
         max_iter=40, max_tridiag_iter=12, tolerance=1e-7)

    # (e) early convergence: identity-like matrix, tridiag update switches itself off (:326)
    a = torch.eye(20, dtype=torch.float64) * 2.0
    b = torch.randn(20, 3, dtype=torch.float64, generator=g)
    x, t = linear_cg(a.matmul, b, n_tridiag=3, max_iter=15, max_tridiag_iter=8, tolerance=1e-3)
    save("cg_identity_f64", A=npy(a), rhs=npy(b), x=npy(x), t_mat=npy(t), n_tridiag=3, max_iter=15,
         max_tridiag_iter=8, tolerance=1e-3)


def pivchol_cases():
    g = torch.Generator().manual_seed(5)
    for dt, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        k = rbf(40, dt, batch=(2, 3), gen=g)
        op = DenseLinearOperator(k)
        L, perm = lo.functions._pivoted_cholesky.PivotedCholesky.apply(op.representation_tree(), 8, None, k)
        save("pivchol_rbf_" + tag, A=npy(k), L=npy(L), perm=npy(perm), rank=8, tol=1e-3)
    # early stop: rank-3 matrix, tolerance reached after 3 steps
    w = torch.randn(25, 3, dtype=torch.float64, generator=g)
    k = w @ w.mT
    L, perm = lo.functions._pivoted_cholesky.PivotedCholesky.apply(DenseLinearOperator(k).representation_tree(), 10,
                                                                   1e-6, k)
    save("pivchol_lowrank_f64", A=npy(k), L=npy(L), perm=npy(perm), rank=10, tol=1e-6)


def precond_cases():
    g = torch.Generator().manual_seed(7)
    for tag, const in (("const", True), ("varying", False)):
        k = rbf(50, torch.float64, batch=(2,), gen=g)
        d = torch.full((2, 50), 0.3, dtype=torch.float64) if const else 0.1 + torch.rand(2, 50, dtype=torch.float64,
                                                                                           generator=g)
        op = AddedDiagLinearOperator(DenseLinearOperator(k), DiagLinearOperator(d))
        with settings.min_preconditioning_size(4), settings.max_preconditioner_size(6):
            closure, plt, logdet_p = op._preconditioner()
        v = torch.randn(2, 50, 3, dtype=torch.float64, generator=g)
        save("precond_" + tag + "_f64", A=npy(k), d=npy(d), L=npy(op._piv_chol_self), v=npy(v), minv_v=npy(closure(v)),
             logdet_p=npy(logdet_p), rank=6)


class _Injected(AddedDiagLinearOperator):
    probes = None

    def _probe_vectors_and_norms(self):
        return self.probes, torch.ones_like(self.probes[..., :1, :])


def iqld_cases():
    g = torch.Generator().manual_seed(3)
    for tag, dt, batch, precond in (
        ("noprecond_f64", torch.float64, (3,), False),
        ("precond_f64", torch.float64, (2,), True),
        ("precond_f32", torch.float32, (2, 2), True),
        ("noprecond_f32", torch.float32, (), False),
    ):
        n, s = 60, 8
        k = wishart(n, dt, batch=batch, rank=20, gen=g)
        d = torch.full(batch + (n,), 0.5, dtype=dt)
        rhs = torch.randn(*batch, n, 2, dtype=dt, generator=g)
        probes = torch.randn(*batch, n, s, dtype=dt, generator=g)
        probes = probes / probes.norm(dim=-2, keepdim=True)
        op = _Injected(DenseLinearOperator(k), DiagLinearOperator(d))
        op.probes = probes
        with settings.max_cholesky_size(0), settings.min_preconditioning_size(4 if precond else 10**6), \
                settings.max_preconditioner_size(6):
            iq, ld = op.inv_quad_logdet(rhs, logdet=True)
            iq_nr, _ = op.inv_quad_logdet(rhs, logdet=True, reduce_inv_quad=False)
        save("iqld_dense_" + tag, A=npy(k), d=npy(d), rhs=npy(rhs), probes=npy(probes), inv_quad=npy(iq),
             inv_quad_noreduce=npy(iq_nr), logdet=npy(ld), precond=int(precond), rank=6)

    # RNG-drawn probes: replicate the draw order (root samples first, then diag samples)
    n, s, dt, batch = 60, 5, torch.float64, (2,)
    k = wishart(n, dt, batch=batch, rank=20, gen=g)
    d = 0.2 + torch.rand(*batch, n, dtype=dt, generator=g)
    rhs = torch.randn(*batch, n, 1, dtype=dt, generator=g)
    op = AddedDiagLinearOperator(DenseLinearOperator(k), DiagLinearOperator(d))
    with settings.max_cholesky_size(0), settings.min_preconditioning_size(4), settings.max_preconditioner_size(6), \
            settings.num_trace_samples(s):
        torch.manual_seed(1234)
        iq, ld = op.inv_quad_logdet(rhs, logdet=True)
    torch.manual_seed(1234)
    eps_root = torch.randn(*batch, 6, s, dtype=dt)
    eps_diag = torch.randn(s, *batch, n, dtype=dt)
    save("iqld_dense_rng_f64", A=npy(k), d=npy(d), rhs=npy(rhs), eps_root=npy(eps_root), eps_diag=npy(eps_diag),
         inv_quad=npy(iq), logdet=npy(ld), rank=6)

    # identity precond_lt probes (no preconditioner): randn(S, *b, N)
    op = AddedDiagLinearOperator(DenseLinearOperator(k), DiagLinearOperator(d))
    with settings.max_cholesky_size(0), settings.num_trace_samples(s):
        torch.manual_seed(4321)
        iq, ld = op.inv_quad_logdet(rhs, logdet=True)
    torch.manual_seed(4321)
    eps_diag = torch.randn(s, *batch, n, dtype=dt)
    save("iqld_dense_rng_noprecond_f64", A=npy(k), d=npy(d), rhs=npy(rhs), eps_diag=npy(eps_diag), inv_quad=npy(iq),
         logdet=npy(ld))


def slq_cases():
    g = torch.Generator().manual_seed(21)
    a = wishart(40, torch.float64, batch=(3,), gen=g, jitter=0.3)
    b = torch.randn(3, 40, 6, dtype=torch.float64, generator=g)
    _, t = linear_cg(a.matmul, b, n_tridiag=6, max_iter=40, max_tridiag_iter=12, tolerance=1e-9)
    t[0, 0] = -t[0, 0]  # force some negative eigenvalues through the masking branch (lanczos.py:184-187)
    evals, evecs = lanczos_tridiag_to_diag(t.clone())
    (ld,) = StochasticLQ().to_dense(torch.Size((40, 40)), evals, evecs, [lambda x: x.log()])
    save("slq_f64", t_mat=npy(t), evals=npy(evals), evecs=npy(evecs), logdet=npy(ld), n=40)


def structured_cases():
    g = torch.Generator().manual_seed(9)
    dt = torch.float64
    # Kronecker: matmul, diag, rows via __getitem__
    fs = [wishart(m, dt, batch=(2,), gen=g, jitter=0.1) for m in (3, 4, 5)]
    op = KroneckerProductLinearOperator(*fs)
    x = torch.randn(2, 60, 4, dtype=dt, generator=g)
    idx = torch.tensor([7, 58])
    rows = torch.stack([op[0, 7].to_dense() if hasattr(op[0, 7], "to_dense") else op[0, 7],
                        op[1, 58].to_dense() if hasattr(op[1, 58], "to_dense") else op[1, 58]])
    save("kron_f64", f0=npy(fs[0]), f1=npy(fs[1]), f2=npy(fs[2]), x=npy(x), y=npy(op._matmul(x)),
         diag=npy(op._diagonal()), row_idx=npy(idx), rows=npy(rows), dense=npy(op.to_dense()))
    # Toeplitz
    for tag, tdt in (("f64", torch.float64), ("f32", torch.float32)):
        col = torch.exp(-0.5 * (torch.arange(37, dtype=tdt) / 5.0) ** 2).repeat(2, 1) * torch.tensor([[1.0], [2.0]],
                                                                                                      dtype=tdt)
        op = ToeplitzLinearOperator(col)
        x = torch.randn(2, 37, 3, dtype=tdt, generator=g)
        save("toeplitz_" + tag, col=npy(col), x=npy(x), y=npy(op._matmul(x)), dense=npy(op.to_dense()))
    # low-rank root + diag (Woodbury direct path)
    u = torch.randn(2, 50, 5, dtype=dt, generator=g) / 3
    d = 0.3 + torch.rand(2, 50, dtype=dt, generator=g)
    rhs = torch.randn(2, 50, 3, dtype=dt, generator=g)
    op = LowRankRootLinearOperator(u) + DiagLinearOperator(d)
    assert type(op).__name__ == "LowRankRootAddedDiagLinearOperator"
    iq, ld = op.inv_quad_logdet(rhs, logdet=True)
    save("lowrank_f64", U=npy(u), d=npy(d), rhs=npy(rhs), solve=npy(op.solve(rhs)), inv_quad=npy(iq), logdet=npy(ld))
    # Lanczos with re-orthogonalisation
    a = wishart(30, dt, batch=(2,), gen=g, jitter=0.2)
    init = torch.randn(2, 30, 3, dtype=dt, generator=g)
    q, t = lanczos_tridiag(a.matmul, 8, dtype=dt, device=a.device, matrix_shape=a.shape[-2:], batch_shape=a.shape[:-2],
                           init_vecs=init)
    save("lanczos_f64", A=npy(a), init=npy(init), q_mat=npy(q), t_mat=npy(t), max_iter=8)
    # Kronecker + diag through CG (AddedDiag built directly), injected probes
    fs = [wishart(m, dt, gen=g, jitter=0.1) for m in (4, 5, 6)]
    d = torch.full((120,), 0.5, dtype=dt)
    probes = torch.randn(120, 6, dtype=dt, generator=g)
    probes = probes / probes.norm(dim=-2, keepdim=True)
    rhs = torch.randn(120, 1, dtype=dt, generator=g)
    op = _Injected(KroneckerProductLinearOperator(*fs), DiagLinearOperator(d))
    op.probes = probes
    with settings.max_cholesky_size(0), settings.min_preconditioning_size(4), settings.max_preconditioner_size(5):
        iq, ld = op.inv_quad_logdet(rhs, logdet=True)
    save("iqld_kron_f64", f0=npy(fs[0]), f1=npy(fs[1]), f2=npy(fs[2]), d=npy(d), rhs=npy(rhs), probes=npy(probes),
         inv_quad=npy(iq), logdet=npy(ld), rank=5, L=npy(op._piv_chol_self))
    # Toeplitz + diag through CG
    col = torch.exp(-0.5 * (torch.arange(80, dtype=dt) / 4.0) ** 2)
    d = torch.full((80,), 0.5, dtype=dt)
    probes = torch.randn(80, 6, dtype=dt, generator=g)
    probes = probes / probes.norm(dim=-2, keepdim=True)
    rhs = torch.randn(80, 1, dtype=dt, generator=g)
    op = _Injected(ToeplitzLinearOperator(col), DiagLinearOperator(d))
    op.probes = probes
    with settings.max_cholesky_size(0), settings.min_preconditioning_size(4), settings.max_preconditioner_size(5):
        iq, ld = op.inv_quad_logdet(rhs, logdet=True)
    save("iqld_toeplitz_f64", col=npy(col), d=npy(d), rhs=npy(rhs), probes=npy(probes), inv_quad=npy(iq), logdet=npy(ld),
         rank=5, L=npy(op._piv_chol_self))


if __name__ == "__main__":
    torch.set_num_threads(4)
    cg_cases()
    pivchol_cases()
    precond_cases()
    iqld_cases()
    slq_cases()
    structured_cases()

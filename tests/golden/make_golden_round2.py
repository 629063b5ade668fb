#!/usr/bin/env python3
"""Round-2 fixtures, produced by importing and RUNNING THE REFERENCE (read-only at /root/reference) in the build
container:

    cd /tmp && python /root/repo/tests/golden/make_golden_round2.py

* BASELINE configs[0] at its real shape (N = 512, fp64, 16 probes, ``max_cholesky_size(0)``);
* Kronecker in fp32 (small batched + 20x20x20), Kronecker-through-CG in fp32;
* pivot permutations of Kronecker / Toeplitz operators (bit-exact index work);
* Toeplitz at N = 4096 in fp32 / fp64;
* generic-operator pivoted Cholesky (RootLinearOperator, Dense + Root sum) and ``AddedDiag(Root, Diag)`` through the
  preconditioned path;
* backward passes: ``solve`` (with / without left tensor), ``inv_quad``, Toeplitz ``inv_quad_logdet``.
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
    RootLinearOperator,
    ToeplitzLinearOperator,
)

OUT = os.path.dirname(os.path.abspath(__file__))
warnings.simplefilter("ignore")


def npy(t):
    return t.detach().cpu().numpy()


def save(name, **arrs):
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **{k: np.asarray(v) for k, v in arrs.items()})
    print("wrote", name, {k: np.asarray(v).shape for k, v in arrs.items()})


def wishart(n, dtype, batch=(), rank=None, gen=None, jitter=0.0):
    rank = rank or n
    w = torch.randn(*batch, n, rank, dtype=dtype, generator=gen)
    return w @ w.mT / rank + jitter * torch.eye(n, dtype=dtype)


class _Injected(AddedDiagLinearOperator):
    probes = None

    def _probe_vectors_and_norms(self):
        return self.probes, torch.ones_like(self.probes[..., :1, :])


def unit_probes(*shape, dtype, gen):
    p = torch.randn(*shape, dtype=dtype, generator=gen)
    return p / p.norm(dim=-2, keepdim=True)


def cfg1():
    """BASELINE configs[0]: Dense + AddedDiag, N = 512, batch 1, fp64, 16 probes (SURVEY 8d generator)."""
    g = torch.Generator().manual_seed(1234)
    n, s, dt = 512, 16, torch.float64
    w = torch.randn(n, 256, dtype=dt, generator=g)
    sc = torch.logspace(0, -1.5, 256, dtype=dt)
    w = w * sc / sc.norm()
    k = w @ w.mT
    d = torch.full((n,), 0.5, dtype=dt)
    rhs = torch.randn(n, 1, dtype=dt, generator=g)
    probes = unit_probes(n, s, dtype=dt, gen=g)
    op = _Injected(DenseLinearOperator(k), DiagLinearOperator(d))
    op.probes = probes
    with settings.max_cholesky_size(0), settings.num_trace_samples(s):
        iq, ld = op.inv_quad_logdet(rhs, logdet=True)
        sol = op.solve(rhs)
    save("cfg1_dense_f64", W=npy(w), d=npy(d), rhs=npy(rhs), probes=npy(probes), inv_quad=npy(iq), logdet=npy(ld),
         solve=npy(sol))


def kron_cases():
    g = torch.Generator().manual_seed(31)
    dt = torch.float32
    fs = [wishart(m, dt, batch=(2,), gen=g, jitter=0.1) for m in (6, 7, 8)]
    op = KroneckerProductLinearOperator(*fs)
    x = torch.randn(2, 336, 4, dtype=dt, generator=g)
    save("kron_f32", f0=npy(fs[0]), f1=npy(fs[1]), f2=npy(fs[2]), x=npy(x), y=npy(op._matmul(x)),
         diag=npy(op._diagonal()))
    fs = [wishart(20, dt, gen=g, jitter=0.1) for _ in range(3)]
    op = KroneckerProductLinearOperator(*fs)
    x = torch.randn(8000, 5, dtype=dt, generator=g)
    save("kron_mid_f32", f0=npy(fs[0]), f1=npy(fs[1]), f2=npy(fs[2]), x=npy(x), y=npy(op._matmul(x)))
    # Kronecker + diag through preconditioned CG in fp32, injected probes, pivots recorded
    fs = [wishart(m, dt, gen=g, jitter=0.1) for m in (8, 9, 10)]
    n = 720
    d = torch.full((n,), 0.5, dtype=dt)
    probes = unit_probes(n, 6, dtype=dt, gen=g)
    rhs = torch.randn(n, 1, dtype=dt, generator=g)
    kron = KroneckerProductLinearOperator(*fs)
    op = _Injected(kron, DiagLinearOperator(d))
    op.probes = probes
    with settings.max_cholesky_size(0), settings.min_preconditioning_size(4), settings.max_preconditioner_size(5):
        iq, ld = op.inv_quad_logdet(rhs, logdet=True)
        L, perm = kron.pivoted_cholesky(rank=5, return_pivots=True)
    save("iqld_kron_f32", f0=npy(fs[0]), f1=npy(fs[1]), f2=npy(fs[2]), d=npy(d), rhs=npy(rhs), probes=npy(probes),
         inv_quad=npy(iq), logdet=npy(ld), rank=5, L=npy(L), perm=npy(perm))


def pivot_cases():
    g = torch.Generator().manual_seed(41)
    dt = torch.float64
    fs = [wishart(m, dt, batch=(2,), gen=g, jitter=0.1) for m in (4, 5, 6)]
    L, perm = KroneckerProductLinearOperator(*fs).pivoted_cholesky(rank=8, return_pivots=True)
    save("pivchol_kron_f64", f0=npy(fs[0]), f1=npy(fs[1]), f2=npy(fs[2]), L=npy(L), perm=npy(perm), rank=8)
    col = torch.exp(-0.5 * (torch.arange(80, dtype=dt) / 4.0) ** 2).repeat(2, 1) * torch.tensor([[1.0], [1.7]], dtype=dt)
    L, perm = ToeplitzLinearOperator(col).pivoted_cholesky(rank=8, return_pivots=True)
    save("pivchol_toeplitz_f64", col=npy(col), L=npy(L), perm=npy(perm), rank=8)
    # generic operators: rows through _get_indices
    u = torch.randn(2, 70, 12, dtype=dt, generator=g) / 3
    L, perm = RootLinearOperator(u).pivoted_cholesky(rank=8, return_pivots=True)
    save("pivchol_root_f64", U=npy(u), L=npy(L), perm=npy(perm), rank=8)
    a = wishart(70, dt, batch=(2,), rank=9, gen=g)
    L, perm = (DenseLinearOperator(a) + RootLinearOperator(u)).pivoted_cholesky(rank=10, return_pivots=True)
    save("pivchol_sum_f64", A=npy(a), U=npy(u), L=npy(L), perm=npy(perm), rank=10)
    # AddedDiag(Root(U), Diag) through the preconditioned path (crashes without the generic driver)
    d = 0.3 + torch.rand(2, 70, dtype=dt, generator=g)
    probes = unit_probes(2, 70, 6, dtype=dt, gen=g)
    rhs = torch.randn(2, 70, 2, dtype=dt, generator=g)
    op = _Injected(RootLinearOperator(u), DiagLinearOperator(d))
    op.probes = probes
    with settings.max_cholesky_size(0), settings.min_preconditioning_size(4), settings.max_preconditioner_size(6):
        iq, ld = op.inv_quad_logdet(rhs, logdet=True)
    save("iqld_root_f64", U=npy(u), d=npy(d), rhs=npy(rhs), probes=npy(probes), inv_quad=npy(iq), logdet=npy(ld),
         rank=6, L=npy(op._piv_chol_self))


def toeplitz_cases():
    g = torch.Generator().manual_seed(51)
    for tag, dt in (("f64", torch.float64), ("f32", torch.float32)):
        n = 4096
        col = torch.exp(-0.5 * (torch.arange(n, dtype=dt) / 50.0) ** 2).repeat(2, 1) * torch.tensor([[1.0], [1.5]], dtype=dt)
        op = ToeplitzLinearOperator(col)
        x = torch.randn(2, n, 3, dtype=dt, generator=g)
        save("toeplitz_big_" + tag, col=npy(col), x=npy(x), y=npy(op._matmul(x)))


def backward_cases():
    g = torch.Generator().manual_seed(61)
    dt = torch.float64
    n = 60
    # solve: no left tensor / left tensor, preconditioned, tight tolerance
    for tag, has_left in (("solve", False), ("solve_left", True)):
        a = wishart(n, dt, batch=(2,), rank=20, gen=g).requires_grad_(True)
        d = (0.3 + torch.rand(2, n, dtype=dt, generator=g)).requires_grad_(True)
        rhs = torch.randn(2, n, 3, dtype=dt, generator=g).requires_grad_(True)
        lhs = torch.randn(2, 4, n, dtype=dt, generator=g).requires_grad_(True)
        w = torch.randn(2, 4 if has_left else n, 3, dtype=dt, generator=g)
        op = AddedDiagLinearOperator(DenseLinearOperator(a), DiagLinearOperator(d))
        with settings.max_cholesky_size(0), settings.min_preconditioning_size(4), settings.max_preconditioner_size(6), \
                settings.cg_tolerance(1e-10), settings.max_cg_iterations(300):
            res = op.solve(rhs, lhs) if has_left else op.solve(rhs)
            (res * w).sum().backward()
        save("backward_" + tag + "_f64", A=npy(a), d=npy(d), rhs=npy(rhs), lhs=npy(lhs), w=npy(w), res=npy(res),
             grad_A=npy(a.grad), grad_d=npy(d.grad), grad_rhs=npy(rhs.grad),
             grad_lhs=npy(lhs.grad) if has_left else np.zeros(0), rank=6, tol=1e-10)
    # inv_quad
    a = wishart(n, dt, batch=(2,), rank=20, gen=g).requires_grad_(True)
    d = (0.3 + torch.rand(2, n, dtype=dt, generator=g)).requires_grad_(True)
    rhs = torch.randn(2, n, 3, dtype=dt, generator=g).requires_grad_(True)
    w = torch.randn(2, 3, dtype=dt, generator=g)
    op = AddedDiagLinearOperator(DenseLinearOperator(a), DiagLinearOperator(d))
    with settings.max_cholesky_size(0), settings.min_preconditioning_size(4), settings.max_preconditioner_size(6), \
            settings.cg_tolerance(1e-10), settings.max_cg_iterations(300):
        res = op.inv_quad(rhs, reduce_inv_quad=False)
        (res * w).sum().backward()
    save("backward_inv_quad_f64", A=npy(a), d=npy(d), rhs=npy(rhs), w=npy(w), res=npy(res), grad_A=npy(a.grad),
         grad_d=npy(d.grad), grad_rhs=npy(rhs.grad), rank=6, tol=1e-10)
    # Toeplitz + diag: inv_quad_logdet backward without preconditioner (the column's gradient comes from
    # sym_toeplitz_derivative_quadratic_form, utils/toeplitz.py:164-204)
    n = 80
    col = torch.exp(-0.5 * (torch.arange(n, dtype=dt) / 4.0) ** 2).repeat(2, 1) * torch.tensor([[1.0], [1.4]], dtype=dt)
    col = col.requires_grad_(True)
    d = (0.4 + torch.rand(2, n, dtype=dt, generator=g)).requires_grad_(True)
    rhs = torch.randn(2, n, 2, dtype=dt, generator=g).requires_grad_(True)
    probes = unit_probes(2, n, 6, dtype=dt, gen=g)
    w_iq = torch.randn(2, dtype=dt, generator=g)
    w_ld = torch.randn(2, dtype=dt, generator=g)
    op = _Injected(ToeplitzLinearOperator(col), DiagLinearOperator(d))
    op.probes = probes
    with settings.max_cholesky_size(0), settings.max_preconditioner_size(0):
        iq, ld = op.inv_quad_logdet(rhs, logdet=True)
        ((iq * w_iq).sum() + (ld * w_ld).sum()).backward()
    save("backward_toeplitz_f64", col=npy(col), d=npy(d), rhs=npy(rhs), probes=npy(probes), w_iq=npy(w_iq),
         w_ld=npy(w_ld), inv_quad=npy(iq), logdet=npy(ld), grad_col=npy(col.grad), grad_d=npy(d.grad),
         grad_rhs=npy(rhs.grad))


def backward_constant_diag_cases():
    """Constant diagonals: the reference keeps ONE noise value (added_diag_linear_operator.py:161), so the gradient of
    logdet_P w.r.t. a (B, N) all-equal diagonal lands on element 0 only; with a ConstantDiagLinearOperator it lands on
    diag_values."""
    from linear_operator.operators import ConstantDiagLinearOperator

    g = torch.Generator().manual_seed(71)
    dt = torch.float64
    n, s = 60, 6
    for tag in ("full", "constop"):
        w = torch.randn(2, n, 20, dtype=dt, generator=g)
        k = (w @ w.mT / 20).requires_grad_(True)
        rhs = torch.randn(2, n, 2, dtype=dt, generator=g).requires_grad_(True)
        probes = unit_probes(2, n, s, dtype=dt, gen=g)
        w_iq = torch.randn(2, dtype=dt, generator=g)
        w_ld = torch.randn(2, dtype=dt, generator=g)
        if tag == "full":
            d = torch.full((2, n), 0.4, dtype=dt).requires_grad_(True)
            diag_op = DiagLinearOperator(d)
        else:
            d = torch.tensor([[0.4], [0.7]], dtype=dt).requires_grad_(True)
            diag_op = ConstantDiagLinearOperator(d, diag_shape=n)
        op = _Injected(DenseLinearOperator(k), diag_op)
        op.probes = probes
        with settings.max_cholesky_size(0), settings.min_preconditioning_size(4), settings.max_preconditioner_size(6), \
                settings.cg_tolerance(1e-10), settings.max_cg_iterations(400):
            iq, ld = op.inv_quad_logdet(rhs, logdet=True)
            (w_iq * iq + w_ld * ld).sum().backward()
        save("backward_dense_const_" + tag + "_f64", A=npy(k), d=npy(d), rhs=npy(rhs), probes=npy(probes), w_iq=npy(w_iq),
             w_ld=npy(w_ld), inv_quad=npy(iq), logdet=npy(ld), grad_A=npy(k.grad), grad_d=npy(d.grad),
             grad_rhs=npy(rhs.grad), rank=6)


if __name__ == "__main__":
    torch.set_num_threads(4)
    cfg1()
    kron_cases()
    pivot_cases()
    toeplitz_cases()
    backward_cases()
    backward_constant_diag_cases()

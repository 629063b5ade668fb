"""
oracle/krylov_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A CPU (numpy) restatement of the batched Krylov hot path of cornellius-gp/linear_operator
(`/root/reference`, pure Python on top of PyTorch): modified batched conjugate gradients with
tridiagonal recovery, stochastic Lanczos quadrature, the pivoted-Cholesky preconditioner of
``AddedDiagLinearOperator`` and the structured matmuls that feed them.

Who may use this file: ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` -- as the *checker* or the *timed CPU baseline*, never as a
compute path of ``linear_operator_b200`` (the product fails loudly without its CUDA library).

Parity pin: every function below is checked in ``tests/test_oracle_golden.py`` against fixtures that were
produced by importing and running the reference itself in the build container
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``).  The reference ships no golden vectors of its
own: its tests recompute dense answers at run time, so "reference outputs on seeded
inputs" is the pin.

All ``file:line`` citations are relative to ``/root/reference/linear_operator``.
Layout convention (same as the reference): vectors are ``(*batch, N, C)`` with C fastest.
"""
from __future__ import annotations

import numpy as np

# ------------------------------------------------------------------------------------------------
# defaults of the settings the path reads (settings.py:216,383,405,394,417,453,496,484)
# ------------------------------------------------------------------------------------------------
DEFAULTS = dict(
    cg_tolerance=1.0,
    max_cg_iterations=1000,
    max_lanczos_quadrature_iterations=20,
    max_cholesky_size=800,
    max_preconditioner_size=15,
    min_preconditioning_size=2000,
    preconditioner_tolerance=1e-3,
    num_trace_samples=10,
)


def _col_norm(v):
    return np.sqrt(np.sum(v * v, axis=-2, keepdims=True))


# ------------------------------------------------------------------------------------------------
# modified batched CG  (utils/linear_cg.py:98-359)
# ------------------------------------------------------------------------------------------------
def linear_cg(
    matmul_closure,
    rhs,
    n_tridiag=0,
    tolerance=None,
    eps=1e-10,
    stop_updating_after=1e-10,
    max_iter=None,
    max_tridiag_iter=None,
    initial_guess=None,
    preconditioner=None,
    terminate_cg_by_size=False,
    info=None,
):
    """mBCG: solves A X = RHS for all batch elements and columns at once and (optionally) recovers the
    Lanczos tridiagonals of the first ``n_tridiag`` columns from the CG coefficients.

    Follows utils/linear_cg.py:134-359 step by step (setup :134-242, loop :245-332, epilogue :335-359).
    ``info`` (optional dict) receives ``iterations`` and ``tolerance_reached``.
    Returns ``x`` or ``(x, t_mat)`` with ``t_mat`` shaped ``(n_tridiag, *batch, T, T)``.
    """
    rhs = np.asarray(rhs)
    is_vector = rhs.ndim == 1  # :134-136
    if is_vector:
        rhs = rhs[:, None]
    if max_iter is None:
        max_iter = DEFAULTS["max_cg_iterations"]
    if max_tridiag_iter is None:
        max_tridiag_iter = DEFAULTS["max_lanczos_quadrature_iterations"]
    if tolerance is None:
        tolerance = DEFAULTS["cg_tolerance"]
    if initial_guess is None:
        initial_guess = np.zeros_like(rhs)
    elif initial_guess.ndim == 1:
        initial_guess = initial_guess[:, None]
    precond = preconditioner is not None
    if not precond:
        preconditioner = lambda v: v.copy()  # noqa: E731  (:12-13)
    if max_tridiag_iter > max_iter:  # :159-160
        raise RuntimeError("Getting a tridiagonalization larger than the number of CG iterations run is not possible!")
    if isinstance(matmul_closure, np.ndarray):  # :163-166
        _mat = matmul_closure
        matmul_closure = lambda v: _mat @ v  # noqa: E731
    elif not callable(matmul_closure):
        raise RuntimeError("matmul_closure must be a tensor, or a callable object!")

    dt = rhs.dtype
    num_rows = rhs.shape[-2]
    n_iter = min(max_iter, num_rows) if terminate_cg_by_size else max_iter  # :170
    n_tridiag_iter = min(max_tridiag_iter, num_rows)  # :171
    eps = np.asarray(eps, dtype=dt)  # :172
    sua = np.asarray(stop_updating_after, dtype=dt)

    # column normalisation (:177-183)
    rhs_norm = _col_norm(rhs)
    rhs_is_zero = rhs_norm < eps
    rhs_norm = np.where(rhs_is_zero, np.asarray(1, dt), rhs_norm)
    rhs = rhs / rhs_norm
    initial_guess = initial_guess / rhs_norm

    residual = rhs - matmul_closure(initial_guess)  # :186
    batch_shape = residual.shape[:-2]
    result = np.broadcast_to(initial_guess, residual.shape).copy()  # :190
    if np.isnan(residual).any():  # :199-200
        raise RuntimeError("NaNs encountered when trying to perform matrix-vector multiplication")

    residual_norm = _col_norm(residual)  # :204-205
    has_converged = residual_norm < sua
    ncol = rhs.shape[-1]

    if has_converged.all() and not n_tridiag:  # :207-208
        n_iter = 0
    else:
        precond_residual = preconditioner(residual)  # :213-215
        curr_conjugate_vec = precond_residual.copy()
        residual_inner_prod = np.sum(precond_residual * residual, axis=-2, keepdims=True)
        alpha = np.zeros(batch_shape + (1, ncol), dt)
        beta = np.zeros_like(alpha)

    if n_tridiag:  # :224-236
        t_mat = np.zeros((n_tridiag_iter, n_tridiag_iter) + batch_shape + (n_tridiag,), dt)
        prev_alpha_reciprocal = np.zeros(batch_shape + (n_tridiag,), dt)
        prev_beta = np.zeros_like(prev_alpha_reciprocal)

    update_tridiag = True
    last_tridiag_iter = 0
    tolerance_reached = False
    k = -1
    one = np.asarray(1, dt)
    zero = np.asarray(0, dt)

    for k in range(n_iter):  # :245
        mvms = matmul_closure(curr_conjugate_vec)  # :248
        # alpha = <r, z> / <p, A p> with the "denominator < eps -> 0" rule (:250-260 | :64-74)
        alpha = np.sum(curr_conjugate_vec * mvms, axis=-2, keepdims=True)
        is_zero = alpha < eps
        alpha = np.where(is_zero, one, alpha)
        alpha = residual_inner_prod / alpha
        alpha = np.where(is_zero, zero, alpha)
        alpha = np.where(has_converged, zero, alpha)
        residual = residual - alpha * mvms  # :264 | :78
        precond_residual = preconditioner(residual)  # :268 | :82
        # _jit_linear_cg_updates (:31-46)
        result = result + alpha * curr_conjugate_vec
        beta = residual_inner_prod.copy()
        residual_inner_prod = np.sum(residual * precond_residual, axis=-2, keepdims=True)
        is_zero = beta < eps
        beta = np.where(is_zero, one, beta)
        beta = residual_inner_prod / beta
        beta = np.where(is_zero, zero, beta)
        curr_conjugate_vec = curr_conjugate_vec * beta + precond_residual

        residual_norm = _col_norm(residual)  # :298-300
        residual_norm = np.where(rhs_is_zero, zero, residual_norm)
        has_converged = residual_norm < sua

        if (  # :302-308
            k >= min(10, max_iter - 1)
            and bool(residual_norm.mean(dtype=dt) < tolerance)
            and not (n_tridiag and k < min(n_tridiag_iter, max_iter - 1))
        ):
            tolerance_reached = True
            break

        if n_tridiag and k < n_tridiag_iter and update_tridiag:  # :311-332
            alpha_tridiag = alpha[..., 0, :n_tridiag]
            beta_tridiag = beta[..., 0, :n_tridiag]
            a_is_zero = alpha_tridiag == 0
            alpha_reciprocal = one / np.where(a_is_zero, one, alpha_tridiag)
            if k == 0:
                t_mat[k, k] = alpha_reciprocal
            else:
                t_mat[k, k] = alpha_reciprocal + prev_beta * prev_alpha_reciprocal
                off = np.sqrt(prev_beta) * prev_alpha_reciprocal
                t_mat[k, k - 1] = off
                t_mat[k - 1, k] = off
                if t_mat[k - 1, k].max() < 1e-6:
                    update_tridiag = False
            last_tridiag_iter = k
            prev_alpha_reciprocal = alpha_reciprocal.copy()
            prev_beta = beta_tridiag.copy()

    result = result * rhs_norm  # :335
    if info is not None:
        info["iterations"] = k + 1 if n_iter > 0 else 0
        info["tolerance_reached"] = tolerance_reached
        info["residual_norm_mean"] = float(residual_norm.mean())
    if is_vector:
        result = result[..., 0]
    if n_tridiag:  # :352-357
        t = t_mat[: last_tridiag_iter + 1, : last_tridiag_iter + 1]
        nb = len(batch_shape)
        t = np.transpose(t, (t.ndim - 1,) + tuple(range(2, 2 + nb)) + (0, 1))
        return result, np.ascontiguousarray(t)
    return result


# ------------------------------------------------------------------------------------------------
# tridiagonal eigendecomposition + stochastic Lanczos quadrature
# (utils/lanczos.py:167-189, utils/stochastic_lq.py:45-82)
# ------------------------------------------------------------------------------------------------
def lanczos_tridiag_to_diag(t_mat):
    """eigh of every tridiagonal; negative eigenvalues -> 1 with their eigenvector columns zeroed (:184-187)."""
    evals, evecs = np.linalg.eigh(t_mat)
    mask = evals >= 0
    evecs = evecs * mask[..., None, :].astype(evecs.dtype)
    evals = np.where(mask, evals, np.asarray(1, evals.dtype))
    return evals, evecs


def slq_logdet(n, evals, evecs):
    """(N/S) * sum_probes sum_i V[0,i]^2 log(lambda_i)   (stochastic_lq.py:67-82 with funcs=[log])."""
    num_probes = evals.shape[0]
    res = np.zeros(evals.shape[1:-1], evals.dtype)
    for j in range(num_probes):
        first = evecs[j][..., 0, :]
        res = res + (n / float(num_probes)) * np.sum(first * first * np.log(evals[j]), axis=-1)
    return res


# ------------------------------------------------------------------------------------------------
# pivoted Cholesky  (functions/_pivoted_cholesky.py:13-105)
# ------------------------------------------------------------------------------------------------
def pivoted_cholesky(diag, get_rows, rank, error_tol=None):
    """Greedy diagonal-pivoted partial Cholesky of a (batch of) PSD operator(s).

    ``diag``: ``(*batch, N)`` operator diagonal; ``get_rows(pi)``: callable mapping an int64 array ``(*batch,)``
    of row indices to the rows ``(*batch, N)`` (stands in for apply_permutation/__getitem__,
    utils/permutation.py:9-88).  Returns ``L (*batch, N, m)`` and ``perm (*batch, N)`` int64.
    The loop length m is common to the whole batch (global max of the error, :57).
    """
    if error_tol is None:
        error_tol = DEFAULTS["preconditioner_tolerance"]
    diag = np.array(diag, copy=True)  # :30
    batch_shape = diag.shape[:-1]
    n = diag.shape[-1]
    dt = diag.dtype
    max_iter = min(rank, n)  # :33
    L = np.zeros(batch_shape + (max_iter, n), dt)  # :36-42
    orig_error = diag.max(axis=-1)  # :43
    errors = np.abs(diag).sum(axis=-1) / orig_error  # :44
    perm = np.broadcast_to(np.arange(n, dtype=np.int64), batch_shape + (n,)).copy()  # :47-48

    m = 0
    while m == 0 or (m < max_iter and errors.max() > error_tol):  # :57
        pd = np.take_along_axis(diag, perm[..., m:], axis=-1)  # :61
        max_idx = np.argmax(pd, axis=-1)  # first maximal index, like torch.max on CPU (:62)
        max_val = np.take_along_axis(pd, max_idx[..., None], axis=-1)[..., 0]
        max_idx = max_idx + m  # :63
        old_pi_m = perm[..., m].copy()  # :67-70
        perm[..., m] = np.take_along_axis(perm, max_idx[..., None], axis=-1)[..., 0]
        np.put_along_axis(perm, max_idx[..., None], old_pi_m[..., None], axis=-1)
        pi_m = perm[..., m].copy()

        L_m = L[..., m, :]  # view (:73-74)
        np.put_along_axis(L_m, pi_m[..., None], np.sqrt(max_val)[..., None], axis=-1)

        if m + 1 < n:  # :77-95
            row = get_rows(pi_m)
            pi_i = perm[..., m + 1 :]
            L_m_new = np.take_along_axis(row, pi_i, axis=-1).astype(dt)
            if m > 0:
                L_prev = np.take_along_axis(L[..., :m, :], np.broadcast_to(pi_i[..., None, :], batch_shape + (m, pi_i.shape[-1])), axis=-1)
                update = np.take_along_axis(L[..., :m, :], np.broadcast_to(pi_m[..., None, None], batch_shape + (m, 1)), axis=-1)
                L_m_new = L_m_new - np.sum(update * L_prev, axis=-2)
            L_m_new = L_m_new / np.take_along_axis(L_m, pi_m[..., None], axis=-1)
            np.put_along_axis(L_m, pi_i, L_m_new, axis=-1)
            cur = np.take_along_axis(diag, pi_i, axis=-1)
            np.put_along_axis(diag, pi_i, cur - L_m_new**2, axis=-1)
            errors = np.abs(np.take_along_axis(diag, pi_i, axis=-1)).sum(axis=-1) / orig_error  # :98
        m += 1

    return np.ascontiguousarray(np.swapaxes(L[..., :m, :], -1, -2)), perm  # :104


# ------------------------------------------------------------------------------------------------
# AddedDiag preconditioner from the pivoted-Cholesky factor (operators/added_diag_linear_operator.py:95-184)
# ------------------------------------------------------------------------------------------------
def added_diag_preconditioner(L, diag):
    """Returns ``(closure, logdet_P, Q)`` for M = L L^T + diag(d).

    ``L``: ``(*batch, N, k)``; ``diag``: ``(*batch, N)`` materialised diagonal.  "Constant" is decided at run
    time by comparing with the first element (:149-150).
    """
    *batch_shape, n, k = L.shape
    noise = diag[..., None]  # (*b, N, 1)
    constant = bool(np.array_equal(noise, noise[..., :1, :] * np.ones_like(noise)))
    eye = np.broadcast_to(np.eye(k, dtype=L.dtype), tuple(batch_shape) + (k, k))
    if constant:  # :161-172
        noise1 = noise[..., :1, :]
        q, r = np.linalg.qr(np.concatenate([L, np.sqrt(noise1) * eye], axis=-2))
        q = q[..., :n, :]
        logdet = 2 * np.log(np.abs(np.diagonal(r, axis1=-1, axis2=-2))).sum(-1)
        logdet = logdet + (n - k) * np.log(noise1[..., 0, 0])

        def closure(v):  # :135-140
            qqt = q @ (np.swapaxes(q, -1, -2) @ v)
            return (1 / noise1) * (v - qqt)

    else:  # :174-184
        q, r = np.linalg.qr(np.concatenate([L / np.sqrt(noise), eye], axis=-2))
        q = q[..., :n, :] / np.sqrt(noise)
        logdet = 2 * np.log(np.abs(np.diagonal(r, axis1=-1, axis2=-2))).sum(-1)
        logdet = logdet - np.log(1.0 / noise).sum(axis=(-1, -2))

        def closure(v):
            qqt = q @ (np.swapaxes(q, -1, -2) @ v)
            return (v / noise) - qqt

    return closure, logdet.astype(L.dtype), q


def probes_from_base_samples(L, diag, eps_root, eps_diag):
    """Probe vectors z ~ N(0, L L^T + D) from explicit base samples, then column-normalised
    (functions/_inv_quad_logdet.py:107-110; psd_sum_linear_operator.py:15-18;
    operators/_linear_operator.py:2779-2791; diag_linear_operator.py:273-277).

    ``eps_root``: ``(*batch, k, S)`` (drawn first), ``eps_diag``: ``(S, *batch, N)`` (drawn second).
    Returns ``(probes (*batch, N, S), norms (*batch, 1, S))``.
    """
    z_root = L @ eps_root  # (*b, N, S)
    nb = L.ndim - 2
    z_diag = np.transpose(eps_diag * np.sqrt(diag), tuple(range(1, 1 + nb)) + (nb + 1, 0))  # (*b, N, S)
    z = z_root + z_diag
    norms = _col_norm(z)
    return z / norms, norms


# ------------------------------------------------------------------------------------------------
# structured matmuls  
# ------------------------------------------------------------------------------------------------
def dense_added_diag_matmul(A, d, X):
    """A X + d (.) X  (dense_linear_operator.py:60-64, added_diag_linear_operator.py:72-76)."""
    return A @ X + d[..., None] * X


def kron_matmul(factors, X):
    """(K1 (x) K2 (x) ...) X by successive mode products (kronecker_product_linear_operator.py:34-45)."""
    batch_shape = np.broadcast_shapes(X.shape[:-2], *[f.shape[:-2] for f in factors])
    c = X.shape[-1]
    res = np.broadcast_to(X, batch_shape + X.shape[-2:])
    for f in factors:
        ni = f.shape[-1]
        res = res.reshape(batch_shape + (ni, -1))
        fac = f @ res
        fac = fac.reshape(batch_shape + (f.shape[-2], -1, c))
        res = np.swapaxes(fac, -3, -2).reshape(batch_shape + (-1, c))
    return res


def kron_diag(factors):
    """diagonal of a Kronecker product (kronecker_product_linear_operator.py:20-27)."""
    d = np.diagonal(factors[0], axis1=-1, axis2=-2)
    for f in factors[1:]:
        df = np.diagonal(f, axis1=-1, axis2=-2)
        d = (d[..., :, None] * df[..., None, :]).reshape(d.shape[:-1] + (-1,))
    return d


def kron_rows(factors, idx):
    """rows ``idx (*batch,)`` of the Kronecker product (kronecker_product_linear_operator.py:198-216)."""
    sizes = [f.shape[-1] for f in factors]
    n = int(np.prod(sizes))
    cols = np.arange(n, dtype=np.int64)
    out = None
    rem_r = idx
    rem_c = cols
    stride = n
    for f, s in zip(factors, sizes):
        stride //= s
        ri = rem_r // stride
        ci = rem_c // stride
        rem_r = rem_r % stride
        rem_c = rem_c % stride
        fb = np.broadcast_to(f, idx.shape + f.shape[-2:])
        rows = np.take_along_axis(fb, ri[..., None, None].repeat(s, axis=-1), axis=-2)[..., 0, :]  # (*b, s)
        vals = rows[..., ci]
        out = vals if out is None else out * vals
    return out


def sym_toeplitz_matmul(col, X):
    """Symmetric Toeplitz matmul through a length-(2N-1) circulant embedding and complex FFTs
    (utils/toeplitz.py:98-149,152-161)."""
    n = col.shape[-1]
    row = col
    batch_shape = np.broadcast_shapes(col.shape[:-1], X.shape[:-2])
    col_b = np.broadcast_to(col, batch_shape + (n,))
    row_b = np.broadcast_to(row, batch_shape + (n,))
    Xb = np.broadcast_to(X, batch_shape + X.shape[-2:])
    c = np.zeros(batch_shape + (2 * n - 1,), col.dtype)
    c[..., :n] = col_b
    c[..., n:] = row_b[..., 1:][..., ::-1]
    t = np.zeros(batch_shape + (X.shape[-1], 2 * n - 1), X.dtype)
    t[..., :n] = np.swapaxes(Xb, -1, -2)
    cdt = np.complex64 if X.dtype == np.float32 else np.complex128
    fc = np.fft.fft(c).astype(cdt)[..., None, :]
    ft = np.fft.fft(t).astype(cdt)
    out = np.fft.ifft(fc * ft).astype(cdt).real[..., :n]
    return np.ascontiguousarray(np.swapaxes(out, -1, -2)).astype(X.dtype)


def sym_toeplitz_rows(col, idx):
    """rows ``idx`` of a symmetric Toeplitz matrix: T[i,j] = col[|i-j|] (toeplitz_linear_operator.py:38-40)."""
    n = col.shape[-1]
    j = np.arange(n, dtype=np.int64)
    d = np.abs(idx[..., None] - j)
    return np.take_along_axis(np.broadcast_to(col, idx.shape + (n,)), d, axis=-1)


def root_matmul(U, X):
    """U (U^T X)  (root_linear_operator.py:68-72)."""
    return U @ (np.swapaxes(U, -1, -2) @ X)


# ------------------------------------------------------------------------------------------------
# low-rank root + diagonal: Woodbury (operators/low_rank_root_added_diag_linear_operator.py:36-101)
# ------------------------------------------------------------------------------------------------
def lowrank_chol_cap(U, d):
    """chol(I + U^T D^-1 U)  (:36-47)."""
    r = U.shape[-1]
    cap = np.eye(r, dtype=U.dtype) + np.swapaxes(U, -1, -2) @ (U / d[..., None])
    return np.linalg.cholesky(cap)


def lowrank_solve(U, d, rhs):
    """D^-1 b - D^-1 U (I + U^T D^-1 U)^-1 U^T D^-1 b  (:62-87)."""
    chol = lowrank_chol_cap(U, d)
    dinv_b = rhs / d[..., None]
    res = np.swapaxes(U, -1, -2) @ dinv_b
    res = np.linalg.solve(chol, res)
    res = np.linalg.solve(np.swapaxes(chol, -1, -2), res)
    res = (U @ res) / d[..., None]
    return dinv_b - res


def lowrank_logdet(U, d):
    """2 sum log diag chol(cap) + sum log d  (:95-101)."""
    chol = lowrank_chol_cap(U, d)
    return 2 * np.log(np.diagonal(chol, axis1=-1, axis2=-2)).sum(-1) + np.log(d).sum(-1)


def lowrank_inv_quad_logdet(U, d, rhs, reduce_inv_quad=True):
    """(:114-160)"""
    s = lowrank_solve(U, d, rhs)
    iq = (rhs * s).sum(-2)
    if reduce_inv_quad:
        iq = iq.sum(-1)
    return iq, lowrank_logdet(U, d)


# ------------------------------------------------------------------------------------------------
# the whole forward of InvQuadLogdet for a Krylov-path operator
# (operators/_linear_operator.py:1688-1804 + functions/_inv_quad_logdet.py:27-161)
# ------------------------------------------------------------------------------------------------
def inv_quad_logdet(
    matmul,
    n,
    rhs,
    probes,
    preconditioner=None,
    logdet_p=0.0,
    reduce_inv_quad=True,
    tolerance=None,
    max_iter=None,
    max_tridiag_iter=None,
    info=None,
):
    """``probes``: column-normalised ``(*batch, N, S)`` (what ``_probe_vectors_and_norms`` hands in, or what
    functions/_inv_quad_logdet.py:107-110 builds).  ``rhs``: ``(*batch, N, R)`` or None.
    Returns ``(inv_quad, logdet, solves)``."""
    s = probes.shape[-1]
    full_rhs = probes if rhs is None else np.concatenate([probes, rhs], axis=-1)  # :118-132 probes FIRST
    solves, t_mat = linear_cg(
        matmul,
        full_rhs,
        n_tridiag=s,
        tolerance=tolerance,
        max_iter=max_iter,
        max_tridiag_iter=max_tridiag_iter,
        preconditioner=preconditioner,
        info=info,
    )
    if np.isnan(t_mat).any():  # :141-142
        logdet = np.asarray(np.nan, dtype=probes.dtype)
    else:
        evals, evecs = lanczos_tridiag_to_diag(t_mat)
        logdet = slq_logdet(n, evals, evecs)
    inv_quad = None
    if rhs is not None:
        inv_quad = (solves[..., s:] * rhs).sum(-2)  # :151-153
        if reduce_inv_quad:
            inv_quad = inv_quad.sum(-1)
    return inv_quad, logdet + logdet_p, solves


def dense_added_diag_inv_quad_logdet(
    A,
    d,
    rhs,
    probes=None,
    base_samples=None,
    precond_rank=None,
    min_precond_size=None,
    precond_tol=None,
    info=None,
    **cg_kwargs,
):
    """End-to-end restatement for ``AddedDiagLinearOperator(DenseLinearOperator(A), DiagLinearOperator(d))``:
    pivoted-Cholesky preconditioner if N >= min_preconditioning_size (added_diag_linear_operator.py:115-116),
    probes either injected or built from ``base_samples=(eps_root, eps_diag)``, then mBCG + SLQ."""
    n = A.shape[-1]
    if precond_rank is None:
        precond_rank = DEFAULTS["max_preconditioner_size"]
    if min_precond_size is None:
        min_precond_size = DEFAULTS["min_preconditioning_size"]
    matmul = lambda v: dense_added_diag_matmul(A, d, v)  # noqa: E731
    closure, logdet_p, L = None, 0.0, None
    if precond_rank > 0 and n >= min_precond_size:
        batch_shape = A.shape[:-2]

        def get_rows(pi):
            Ab = np.broadcast_to(A, batch_shape + A.shape[-2:])
            return np.take_along_axis(Ab, pi[..., None, None].repeat(n, axis=-1), axis=-2)[..., 0, :]

        L, _ = pivoted_cholesky(np.diagonal(A, axis1=-1, axis2=-2), get_rows, precond_rank, precond_tol)
        if np.isnan(L).any():  # :126-131
            L = None
        else:
            closure, logdet_p, _ = added_diag_preconditioner(L, d)
    if probes is None:
        eps_root, eps_diag = base_samples
        if L is not None:
            probes, _ = probes_from_base_samples(L, d, eps_root, eps_diag)
        else:  # identity precond_lt: plain randn, (S, *b, N) -> (*b, N, S)  (identity_linear_operator.py:262-266)
            nb = A.ndim - 2
            z = np.transpose(eps_diag, tuple(range(1, 1 + nb)) + (nb + 1, 0))
            probes = z / _col_norm(z)
    if info is not None:
        info["precond_rank"] = 0 if L is None else L.shape[-1]
    return inv_quad_logdet(matmul, n, rhs, probes, closure, logdet_p, info=info, **cg_kwargs)


# ------------------------------------------------------------------------------------------------
# Lanczos with full re-orthogonalisation  (utils/lanczos.py:9-164)
# ------------------------------------------------------------------------------------------------
def lanczos_tridiag(matmul_closure, max_iter, init_vecs, tol=1e-5):
    """``init_vecs``: ``(*batch, N, C)``.  Returns ``q_mat (C, *batch, N, T)``, ``t_mat (C, *batch, T, T)``
    (single init vector: leading dim squeezed, :160-162)."""
    batch_shape = init_vecs.shape[:-2]
    n, c = init_vecs.shape[-2:]
    dt = init_vecs.dtype
    num_iter = min(max_iter, n)
    q_mat = np.zeros((num_iter,) + batch_shape + (n, c), dt)
    t_mat = np.zeros((num_iter, num_iter) + batch_shape + (c,), dt)
    q0 = init_vecs / np.sqrt((init_vecs**2).sum(-2))[..., None, :]
    q_mat[0] = q0
    r = matmul_closure(q0)
    a0 = (q0 * r).sum(-2)
    r = r - a0[..., None, :] * q0
    b0 = np.sqrt((r**2).sum(-2))
    t_mat[0, 0] = a0
    t_mat[0, 1] = b0
    t_mat[1, 0] = b0
    q_mat[1] = r / b0[..., None, :]
    k = 0
    for k in range(1, num_iter):
        q_prev = q_mat[k - 1]
        q_curr = q_mat[k]
        beta_prev = t_mat[k, k - 1][..., None, :]
        r = matmul_closure(q_curr) - q_prev * beta_prev
        a = (q_curr * r).sum(-2, keepdims=True)
        t_mat[k, k] = a[..., 0, :]
        if k + 1 < num_iter:
            r = r - a * q_curr
            corr = (r[None] * q_mat[: k + 1]).sum(-2, keepdims=True)
            r = r - (q_mat[: k + 1] * corr).sum(0)
            rn = np.sqrt((r**2).sum(-2, keepdims=True))
            r = r / rn
            b = rn[..., 0, :]
            t_mat[k, k + 1] = b
            t_mat[k + 1, k] = b
            inner = (q_mat[: k + 1] * r[None]).sum(-2)
            could = False
            for _ in range(10):
                if not np.sum(inner > tol):
                    could = True
                    break
                corr = (r[None] * q_mat[: k + 1]).sum(-2, keepdims=True)
                r = r - (q_mat[: k + 1] * corr).sum(0)
                rn2 = np.sqrt((r**2).sum(-2, keepdims=True))
                r = r / rn2
                inner = (q_mat[: k + 1] * r[None]).sum(-2)
            q_mat[k + 1] = r
            if np.sum(np.abs(b) > 1e-6) == 0 or not could:
                break
    num_iter = k + 1
    nb = len(batch_shape)
    q = np.transpose(q_mat[:num_iter], (q_mat.ndim - 1,) + tuple(range(1, 1 + nb)) + (q_mat.ndim - 2, 0))
    t = np.transpose(t_mat[:num_iter, :num_iter], (t_mat.ndim - 1,) + tuple(range(2, 2 + nb)) + (0, 1))
    q = np.ascontiguousarray(q)
    t = np.ascontiguousarray(t)
    if c == 1:
        q, t = q[0], t[0]
    return q, t


# ------------------------------------------------------------------------------------------------
# Backward pass of inv_quad_logdet for AddedDiag(Dense(A), Diag(d))   (SURVEY 8f rank 1 -- oracle only so far)
#   InvQuadLogdet.backward            functions/_inv_quad_logdet.py:163-226
#   _bilinear_derivative              operators/_linear_operator.py:336-393 (autograd of sum(left * (K right)))
#   PivotedCholesky.backward          functions/_pivoted_cholesky.py:106-150
#   d logdet_P / d(L, d)              autograd through the QR of added_diag_linear_operator.py:144-184
# ------------------------------------------------------------------------------------------------
def _cholesky_backward(C, gC):
    """Gradient of A -> chol(A) (lower) for a symmetric A, given the gradient w.r.t. the factor: the formula autograd
    uses, C^-T Phi(C^T gC) C^-1 with Phi = tril with a halved diagonal, symmetrised."""
    P = np.tril(np.swapaxes(C, -1, -2) @ gC)
    idx = np.arange(C.shape[-1])
    P[..., idx, idx] *= 0.5
    Cinv = np.linalg.inv(C)
    S = np.swapaxes(Cinv, -1, -2) @ P @ Cinv
    return 0.5 * (S + np.swapaxes(S, -1, -2))


def pivoted_cholesky_backward(A, perm, m, grad_L):
    """Gradient w.r.t. the dense operator ``A (*b, N, N)`` of its rank-m pivoted Cholesky factor ``L (*b, N, m)``
    (functions/_pivoted_cholesky.py:106-150: L is recomputed as K[pi, pi[:m]] chol(K[pi[:m], pi[:m]])^-T, rows permuted
    back, and differentiated by autograd; restated with the explicit Cholesky / triangular-solve adjoints)."""
    *batch_shape, n, _ = A.shape
    gA = np.zeros_like(A)
    for b in np.ndindex(*batch_shape):
        pi = perm[b]
        short = pi[:m]
        Kpm = A[b][np.ix_(pi, short)]                    # apply_permutation(matrix, full, short)   (:125)
        C = np.linalg.cholesky(Kpm[:m])                  # psd_safe_cholesky(Krows[:m])             (:128)
        Lp = np.concatenate([C, np.linalg.solve(C, Kpm[m:].T).T], axis=0)  # res_pivoted            (:131-137)
        G = grad_L[b][pi]                                # gradient in pivoted row order
        # rows m..N-1:  Lp2 = K2 C^-T   ->  gK2 = G2 C^-1,  gC += -(C^-1 G2^T Lp2)^T restricted to the lower triangle
        Cinv = np.linalg.inv(C)
        G2, L2 = G[m:], Lp[m:]
        gK2 = G2 @ Cinv
        gC = G[:m].copy()
        gC -= np.tril(Cinv.T @ (G2.T @ L2))
        gKmm = _cholesky_backward(C, np.tril(gC))
        gKpm = np.concatenate([gKmm, gK2], axis=0)
        np.add.at(gA[b], np.ix_(pi, short), gKpm)        # indexing backward: scatter-add
    return gA


def dense_added_diag_inv_quad_logdet_backward(A, d, rhs, probes, grad_inv_quad, grad_logdet, precond_rank=None,
                                              min_precond_size=None, precond_tol=None, **cg_kwargs):
    """Gradients ``(grad_A, grad_d, grad_rhs)`` of ``sum_b grad_inv_quad_b * inv_quad_b + grad_logdet_b * logdet_b`` as
    the reference's backward computes them (NOT the exact derivative of the stochastic forward estimate: the probe
    solves stand in for K^-1, functions/_inv_quad_logdet.py:178-215).  ``probes`` are injected, column-normalised
    (norms = 1); ``grad_inv_quad`` is per batch element (reduce_inv_quad=True upstream)."""
    n = A.shape[-1]
    s = probes.shape[-1]
    if precond_rank is None:
        precond_rank = DEFAULTS["max_preconditioner_size"]
    if min_precond_size is None:
        min_precond_size = DEFAULTS["min_preconditioning_size"]
    matmul = lambda v: dense_added_diag_matmul(A, d, v)  # noqa: E731
    closure, logdet_p, L, perm = None, 0.0, None, None
    if precond_rank > 0 and n >= min_precond_size:
        batch_shape = A.shape[:-2]

        def get_rows(pi):
            Ab = np.broadcast_to(A, batch_shape + A.shape[-2:])
            return np.take_along_axis(Ab, pi[..., None, None].repeat(n, axis=-1), axis=-2)[..., 0, :]

        L, perm = pivoted_cholesky(np.diagonal(A, axis1=-1, axis2=-2), get_rows, precond_rank, precond_tol)
        closure, logdet_p, _ = added_diag_preconditioner(L, d)
    _, _, solves = inv_quad_logdet(matmul, n, rhs, probes, closure, logdet_p, **cg_kwargs)

    g_ld = np.asarray(grad_logdet)[..., None, None]
    g_iq = np.asarray(grad_inv_quad)[..., None, None]          # the same weight for every rhs column (reduced sum)
    coef = 1.0 / s                                             # :181
    probe_solves = solves[..., :s] * coef * g_ld               # :182-183 (norms = 1)
    ppv = closure(probes) if closure is not None else probes   # :187-190
    iq_solves = solves[..., s:]                                # :199
    neg = -iq_solves * g_iq                                    # :200
    left = np.concatenate([probe_solves, neg], axis=-1)        # :204
    right = np.concatenate([ppv, iq_solves], axis=-1)          # :205
    grad_A = left @ np.swapaxes(right, -1, -2)                 # _bilinear_derivative of the dense part
    grad_d = (left * right).sum(-1)                            # ... and of the diagonal part
    grad_rhs = -2.0 * neg                                      # :214
    if L is not None:
        # precond_lt = L L^T + D:  _bilinear_derivative(-ppv * coef, ppv * g_ld)   (:209-211)
        u, v = -ppv * coef, ppv * g_ld
        Lt = np.swapaxes(L, -1, -2)
        grad_L = u @ np.swapaxes(Lt @ v, -1, -2) + v @ np.swapaxes(Lt @ u, -1, -2)
        grad_d = grad_d + (u * v).sum(-1)
        # logdet_P = log det(L L^T + D) is added to the estimate (operators/_linear_operator.py:1801) and sits in the
        # autograd graph through the QR of added_diag_linear_operator.py:164/:176
        P = L @ Lt + d[..., None] * np.eye(n, dtype=A.dtype)
        Pinv = np.linalg.inv(P)
        grad_L = grad_L + np.asarray(grad_logdet)[..., None, None] * 2.0 * (Pinv @ L)
        dinv = np.diagonal(Pinv, axis1=-1, axis2=-2)
        if np.array_equal(d, np.broadcast_to(d[..., :1], d.shape)):
            # constant diagonal: the reference keeps ONE noise value, `noise.narrow(-2, 0, 1)`
            # (added_diag_linear_operator.py:161), so d logdet_P / d sigma^2 = tr(P^-1) lands on element 0
            grad_d = grad_d.copy()
            grad_d[..., 0] += np.asarray(grad_logdet) * dinv.sum(-1)
        else:
            grad_d = grad_d + np.asarray(grad_logdet)[..., None] * dinv
        grad_A = grad_A + pivoted_cholesky_backward(A, perm, L.shape[-1], grad_L)
    return grad_A, grad_d, grad_rhs


# ------------------------------------------------------------------------------------------------
# Round-2 additions: generic-operator rows, Solve / InvQuad backward, symmetric-Toeplitz derivative
# ------------------------------------------------------------------------------------------------
def root_rows(U, idx):
    """rows ``idx (*batch,)`` of U U^T through RootLinearOperator._get_indices (root_linear_operator.py:47-58):
    (U[row] * U[col]).sum(-1)."""
    Ub = np.broadcast_to(U, idx.shape + U.shape[-2:])
    left = np.take_along_axis(Ub, idx[..., None, None].repeat(U.shape[-1], axis=-1), axis=-2)  # (*b, 1, r)
    return (left * Ub).sum(-1)


def sym_toeplitz_derivative_quadratic_form(left, right):
    """res[..., i] = sum_j u_j^T (dT/dc_i) v_j for left / right (*b, N, C): ones on the i-th sub- and super-diagonal
    (utils/toeplitz.py:164-204; restated as the two cross-correlations it computes through Toeplitz products)."""
    n = left.shape[-2]
    res = np.zeros(left.shape[:-2] + (n,), left.dtype)
    for i in range(n):
        a = (left[..., : n - i, :] * right[..., i:, :]).sum((-2, -1))
        b = (left[..., i:, :] * right[..., : n - i, :]).sum((-2, -1))
        res[..., i] = a + b if i > 0 else a
    return res


def dense_added_diag_solve_backward(A, d, rhs, grad_out, solve_fn, lhs=None):
    """Solve.backward for AddedDiag(Dense(A), Diag(d)) (functions/_solve.py:70-131).  ``solve_fn(b)`` is the forward
    solve K^-1 b of the same call (preconditioned CG).  Returns (grad_A, grad_d, grad_rhs, grad_lhs)."""
    if lhs is None:
        right_solves = solve_fn(rhs)                                   # saved `solves` (:54)
        left_solves = solve_fn(grad_out)                               # Solve.apply(..., grad_output) (:96)
        grad_lhs = None
    else:
        solves = solve_fn(np.concatenate([np.swapaxes(lhs, -1, -2), rhs], -1))   # :48-49
        nl = lhs.shape[-2]
        right_solves = solves[..., nl:]
        left_solves = solves[..., :nl] @ grad_out                      # :113
        grad_lhs = grad_out @ np.swapaxes(right_solves, -1, -2)        # :116
    L = np.concatenate([left_solves, right_solves], -1)                # :104-107 / :119-122
    R = -0.5 * np.concatenate([right_solves, left_solves], -1)
    grad_A = L @ np.swapaxes(R, -1, -2)                                # dense_linear_operator.py:69-71
    grad_d = (L * R).sum(-1)                                           # diag_linear_operator.py:37-45
    return grad_A, grad_d, left_solves, grad_lhs


def dense_added_diag_inv_quad_backward(A, d, rhs, grad_out, solve_fn):
    """InvQuad.backward (functions/_inv_quad.py:63-93); grad_out (*b, C) weights the un-reduced inverse quadratic
    forms.  Returns (grad_A, grad_d, grad_rhs)."""
    solves = solve_fn(rhs)
    neg = -solves * grad_out[..., None, :]
    grad_A = neg @ np.swapaxes(solves, -1, -2)
    grad_d = (neg * solves).sum(-1)
    return grad_A, grad_d, -2.0 * neg

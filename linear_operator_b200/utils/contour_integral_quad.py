"""Contour-integral quadrature for K^{1/2} b and K^{-1/2} b (reference: utils/contour_integral_quad.py:14-156).

Host logic as in the reference: a 20-step Lanczos run (through linear_cg's tridiagonal) bounds the spectrum, Jacobi
elliptic functions (scipy, on a handful of scalars) give the quadrature shifts and weights, and ONE shifted MINRES run
solves all shifted systems (csrc/minres.cu)."""
from __future__ import annotations

import math
import warnings

import torch

from .. import settings
from . import linear_cg as _cg_module  # noqa: F401  (the package attribute `utils.linear_cg` is the function)
from .minres import minres
from .warnings import NumericalWarning


def contour_integral_quad(linear_op, rhs, inverse=False, weights=None, shifts=None, max_lanczos_iter=20,
                          num_contour_quadrature=None, shift_offset=0):
    """Returns ``(solves, weights, no_shift_solves, shifts)``: sum_q weights_q * solves_q ~ K^{-1/2} rhs (``inverse``)
    or K^{1/2} rhs; same conventions as the reference (:27-38)."""
    import numpy as np
    from scipy.special import ellipj, ellipk

    from .. import utils

    if num_contour_quadrature is None:
        num_contour_quadrature = settings.num_contour_quadrature.value()
    output_batch_shape = torch.broadcast_shapes(linear_op.batch_shape, rhs.shape[:-2])
    preconditioner, preconditioner_lt, _ = linear_op._preconditioner()

    def sqrt_precond_matmul(v):  # :47-52
        if preconditioner_lt is not None:
            s, w, _, _ = contour_integral_quad(preconditioner_lt, v, inverse=False)
            return (s * w).sum(0)
        return v

    rhs = sqrt_precond_matmul(rhs)

    if shifts is None:
        num_extra_dims = max(0, rhs.dim() - linear_op.dim())
        lanczos_init = rhs[(*([0] * num_extra_dims), Ellipsis, slice(None), slice(None, 1))].expand(
            *linear_op.shape[:-1], 1)
        with warnings.catch_warnings(), torch.no_grad():
            warnings.simplefilter("ignore", NumericalWarning)
            _, lanczos_mat = utils.linear_cg(  # :66-75
                lambda v: linear_op._matmul(v), rhs=lanczos_init, n_tridiag=1, max_iter=max_lanczos_iter,
                tolerance=1e-5, max_tridiag_iter=max_lanczos_iter, preconditioner=preconditioner,
            )
            lanczos_mat = lanczos_mat.squeeze(0)
        try:
            approx_eigs = torch.linalg.eigvalsh(lanczos_mat)  # a (<= 20 x 20) matrix per batch element: control path
            if approx_eigs.min() <= 0:
                raise RuntimeError
        except RuntimeError:
            approx_eigs = linear_op._diagonal()
        max_eig = approx_eigs.max(dim=-1)[0]
        min_eig = approx_eigs.min(dim=-1)[0]
        k2 = min_eig / max_eig

        flat_shifts = torch.zeros(num_contour_quadrature + 1, k2.numel(), dtype=k2.dtype, device=k2.device)
        flat_weights = torch.zeros(num_contour_quadrature, k2.numel(), dtype=k2.dtype, device=k2.device)
        for i, (sub_k2, sub_min_eig) in enumerate(zip(k2.flatten().tolist(), min_eig.flatten().tolist())):  # :107-127
            Kp = ellipk(1 - sub_k2)
            nq = num_contour_quadrature
            t = 1j * (np.arange(1, nq + 1) - 0.5) * Kp / nq
            sn, cn, dn, _ = ellipj(np.imag(t), 1 - sub_k2)
            cn = 1.0 / cn
            dn = dn * cn
            sn = 1j * sn * cn
            w = np.sqrt(sub_min_eig) * sn
            w_pow2 = np.real(np.power(w, 2))
            flat_shifts[1:, i].copy_(torch.tensor(w_pow2, dtype=rhs.dtype, device=rhs.device))
            constant = -2 * Kp * np.sqrt(sub_min_eig) / (math.pi * nq)
            flat_weights[:, i].copy_(torch.tensor(cn * dn, dtype=rhs.dtype, device=rhs.device).mul_(constant))
        weights = flat_weights.view(num_contour_quadrature, *k2.shape, 1, 1)
        shifts = flat_shifts.view(num_contour_quadrature + 1, *k2.shape)
        shifts.sub_(shift_offset)
        if k2.shape != output_batch_shape:
            weights = torch.stack([w.expand(*output_batch_shape, 1, 1) for w in weights], 0)
            shifts = torch.stack([s.expand(output_batch_shape) for s in shifts], 0)

    with torch.no_grad():
        solves = minres(lambda v: linear_op._matmul(v), rhs, value=-1, shifts=shifts, preconditioner=preconditioner)
    no_shift_solves = solves[0]
    solves = solves[1:]
    if not inverse:
        solves = torch.stack([linear_op._matmul(s) for s in solves], 0)
    return solves, weights, no_shift_solves, shifts

"""Stochastic Lanczos quadrature (reference: utils/stochastic_lq.py:45-82)."""
from __future__ import annotations

import torch

from .. import _kernels


class StochasticLQ:
    """``to_dense(matrix_shape, eigenvalues, eigenvectors, funcs)`` as in the reference: tr f(A) ~ (N/S) sum_j
    sum_i V_j[0,i]^2 f(lambda_ji).  The log-determinant used by ``inv_quad_logdet`` does not go through this class: it
    is fused with the eigendecomposition (``logdet_from_tridiag``).  ``to_dense`` is kept for API parity and evaluates
    ``funcs`` with torch on the (small) eigenvalue tensors."""

    def __init__(self, max_iter=15, num_random_probes=10):
        self.max_iter = max_iter
        self.num_random_probes = num_random_probes

    def to_dense(self, matrix_shape, eigenvalues, eigenvectors, funcs):
        num_probes = eigenvalues.size(0)
        results = []
        first = eigenvectors[..., 0, :]
        for f in funcs:
            vals = f(eigenvalues)
            results.append((first.pow(2) * vals).sum(-1).sum(0) * (matrix_shape[-1] / float(num_probes)))
        return results

    @staticmethod
    def logdet_from_tridiag(t_mat: torch.Tensor, n: int) -> torch.Tensor:
        """(N/S) sum_probes e1^T log(T) e1 for t_mat (S, *batch, T, T) -> (*batch), one fused device pass."""
        return _kernels.tridiag_eigh_slq(t_mat, n, want_logdet=True)["logdet"]

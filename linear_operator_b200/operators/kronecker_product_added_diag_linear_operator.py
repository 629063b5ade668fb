"""KroneckerProductAddedDiagLinearOperator: ``K1 (x) ... (x) Km + sigma^2 I`` solved in the Kronecker eigenbasis -- what
``KroneckerProductLinearOperator.add_jitter / add_diagonal`` dispatch to (reference:
operators/kronecker_product_added_diag_linear_operator.py:51-224).

Constant diagonal: with K_i = Q_i L_i Q_i^T,  (K + s I)^-1 = Q (L + s)^-1 Q^T  and  log|K + s I| = sum log(L + s)  where
Q = Q_1 (x) ... (x) Q_m and L the Kronecker product of the factor spectra (:154-163, :89-94).  The factor
eigendecompositions are m small dense problems (n_i x n_i, torch.linalg.eigh -- off the hot path like the dense Cholesky
branch); everything of size N -- Q^T b, the spectral scaling, Q (.) -- runs in the Kronecker mode-product kernels, in
double like the reference (settings._linalg_dtype_symeig).  Other diagonals take AddedDiagLinearOperator's CG path."""
from __future__ import annotations

import torch

from .. import _kernels
from .added_diag_linear_operator import AddedDiagLinearOperator
from .diag_linear_operator import ConstantDiagLinearOperator, DiagLinearOperator
from .kronecker_product_linear_operator import KroneckerProductLinearOperator


class KroneckerProductAddedDiagLinearOperator(AddedDiagLinearOperator):
    def __init__(self, *linear_ops, preconditioner_override=None):
        super().__init__(*linear_ops, preconditioner_override=preconditioner_override)
        if not isinstance(self._linear_op, KroneckerProductLinearOperator):
            raise RuntimeError("A KroneckerProductAddedDiagLinearOperator needs a KroneckerProductLinearOperator base!")
        self.linear_op, self.diag_tensor = self._linear_op, self._diag_tensor
        self._diag_is_constant = isinstance(self.diag_tensor, ConstantDiagLinearOperator)
        self._eig_cache = None

    def _factor_eig(self):
        """[(evals_i (*b, n_i), evecs_i (*b, n_i, n_i))] in double and the Kronecker spectrum (*b, N)."""
        if self._eig_cache is None:
            pairs = [torch.linalg.eigh(f.double()) for f in self.linear_op._factor_tensors()]
            lam = pairs[0][0]
            for ev, _ in pairs[1:]:
                lam = (lam.unsqueeze(-1) * ev.unsqueeze(-2)).reshape(*lam.shape[:-1], -1)
            self._eig_cache = (pairs, lam)
        return self._eig_cache

    def _preconditioner(self):  # :130-132: solves do not run CG
        if self._diag_is_constant:
            return None, None, None
        return super()._preconditioner()

    def _solve(self, rhs, preconditioner=None, num_tridiag=0):  # :134-163
        if not self._diag_is_constant or num_tridiag:
            return super()._solve(rhs, preconditioner=preconditioner, num_tridiag=num_tridiag)
        pairs, lam = self._factor_eig()
        sigma = self.diag_tensor.diag_values.double()  # (*b, 1)
        rhs64 = rhs.double()
        qt = [q.mT.contiguous() for _, q in pairs]
        res = _kernels.kron_matmul(qt, rhs64)  # Q^T b
        res = _kernels.scale_rows(res, lam + sigma, "div")  # (L + s)^-1 .
        res = _kernels.kron_matmul([q for _, q in pairs], res)  # Q .
        return res.to(rhs.dtype)

    def _logdet(self):  # :86-94
        if not self._diag_is_constant:
            return super().inv_quad_logdet(logdet=True)[1]
        _, lam = self._factor_eig()
        return torch.log(lam + self.diag_tensor.diag_values.double()).sum(-1).to(self.dtype)

    def inv_quad_logdet(self, inv_quad_rhs=None, logdet=False, reduce_inv_quad=True):  # :67-84
        if not self._diag_is_constant:
            return super().inv_quad_logdet(inv_quad_rhs=inv_quad_rhs, logdet=logdet, reduce_inv_quad=reduce_inv_quad)
        inv_quad_term = None
        if inv_quad_rhs is not None:
            inv_quad_term, _ = super().inv_quad_logdet(inv_quad_rhs=inv_quad_rhs, logdet=False,
                                                       reduce_inv_quad=reduce_inv_quad)
        return inv_quad_term, (self._logdet() if logdet else None)

    def logdet(self):
        return self._logdet()


__all__ = ["KroneckerProductAddedDiagLinearOperator"]

"""LowRankRootAddedDiagLinearOperator: ``U U^T + D`` solved directly with Woodbury -- CG never runs
(reference: operators/low_rank_root_added_diag_linear_operator.py:36-193)."""
from __future__ import annotations

import torch

from .. import _kernels
from ..utils.memoize import cached
from .added_diag_linear_operator import AddedDiagLinearOperator
from .diag_linear_operator import ConstantDiagLinearOperator, DiagLinearOperator
from .root_linear_operator import LowRankRootLinearOperator


class LowRankRootAddedDiagLinearOperator(AddedDiagLinearOperator):
    def __init__(self, *linear_ops, preconditioner_override=None):
        if len(linear_ops) > 2:
            raise RuntimeError("An AddedDiagLinearOperator can only have two components")
        if isinstance(linear_ops[0], DiagLinearOperator) and not isinstance(linear_ops[1], LowRankRootLinearOperator):
            raise RuntimeError(
                "A LowRankRootAddedDiagLinearOperator can only be created with a LowRankLinearOperator base!"
            )
        elif isinstance(linear_ops[1], DiagLinearOperator) and not isinstance(linear_ops[0], LowRankRootLinearOperator):
            raise RuntimeError(
                "A LowRankRootAddedDiagLinearOperator can only be created with a LowRankLinearOperator base!"
            )
        super().__init__(*linear_ops, preconditioner_override=preconditioner_override)

    def _shared_root_constant_diag(self):
        """BASELINE config 5: ONE root U (N, r) shared by the whole batch and a constant diagonal sigma_b per batch
        element.  The reference would materialise D^-1 U as (*batch, N, r) (low_rank_root_added_diag_...py:44 through
        diag_linear_operator.py:211-212: 40 TB at the config's sizes); here U^T D^-1 U = (U^T U) / sigma_b, the r x r
        Gram matrix is computed once, and the two tall products become plain (batch x N) x (N x r) GEMMs that read U
        once for the whole batch -- on the tensor cores through ``_kernels.gemm3x``."""
        return (self._shared_root() is not None and isinstance(self._diag_tensor, ConstantDiagLinearOperator)
                and len(self.batch_shape) > 0)

    def _shared_root(self):
        """The (N, r) root when every batch element uses the same one (un-batched, or a stride-0 expanded view as
        AddedDiagLinearOperator's batch broadcasting produces), else None."""
        U = self._linear_op._root_tensor()
        if U.dim() == 2:
            return U
        if all(st == 0 or sz == 1 for st, sz in zip(U.stride()[:-2], U.shape[:-2])):
            return U[(0,) * (U.dim() - 2)]
        return None

    def _sigma(self):
        return self._diag_tensor.diag_values.expand(*self.batch_shape, 1).reshape(-1)  # (B,)

    def _gram(self):
        """G = U^T D^-1 U in double, shared by ``_solve`` and ``_logdet`` (the reference caches chol(I + G), :36-47)."""
        if getattr(self, "_gram_cache", None) is None and self._shared_root_constant_diag():
            U = self._shared_root()
            k = U.shape[-1]
            # U^T U (r x r, contraction over N) in double: on the tensor cores for fp32 roots, with split-K chunks of at
            # most 1024 rows so that the truncating fp32 accumulate of the tensor core stays below ~5e-6 relative on
            # the all-positive diagonal sums, partial sums added in double; CUDA cores otherwise
            g0 = None
            if U.dtype == torch.float32 and U.shape[0] >= 4096:
                g0 = _kernels.gemm3x(U.unsqueeze(0), U.unsqueeze(0), trans_a=True, out_dtype=torch.float64,
                                     splits=min(4096, -(-U.shape[0] // 1024)))
            if g0 is None:
                g0 = _kernels.tn_matmul(U.unsqueeze(0), U.unsqueeze(0), out_dtype=torch.float64)
            g0 = g0.reshape(1, k, k)
            self._gram_cache = (g0 / self._sigma().double().reshape(-1, 1, 1)).reshape(*self.batch_shape, k, k)
        if getattr(self, "_gram_cache", None) is None:
            U = self._linear_op._root_tensor()
            d = self._diag_tensor._diag
            Us = _kernels.scale_rows(U, d, "div_sqrt")  # D^-1/2 U
            self._gram_cache = _kernels.tn_matmul(Us, Us, out_dtype=torch.float64)
        return self._gram_cache

    def _preconditioner(self):
        return None, None, None

    def _solve_preconditioner(self):
        return None

    def _solve(self, rhs, preconditioner=None, num_tridiag=0):  # :62-87
        U = self._linear_op._root_tensor()
        d = self._diag_tensor._diag
        if self._shared_root_constant_diag() and rhs.shape[-1] == 1 and rhs.shape[:-2] == self.batch_shape:
            # x = (b - U (I + U^T U / s)^-1 U^T b / s) / s with the BATCH as the GEMM row dimension: U is read once per
            # product for the whole batch.  Both tall products run on the tensor cores (csrc/gemm3x.cu, 3xTF32):
            #   W = R U / s      (B x N) (N x r), split-K over N, 1/s folded into the reduction
            #   x = (R - w U^T)/s (B x r) (r x N), the subtraction and the 1/s fused into the epilogue
            U = self._shared_root()
            _kernels.require_cuda(rhs, U)
            n, k = U.shape
            sig = self._sigma()
            inv_sig = sig.reciprocal()
            R = rhs.reshape(-1, n)  # (B, N)
            w = _kernels.gemm3x(R.unsqueeze(0), U.unsqueeze(0), row_alpha=inv_sig.unsqueeze(0))  # U^T D^-1 b
            if w is None:  # fp64 / unaligned: CUDA-core kernels
                w = _kernels.tn_matmul(U.unsqueeze(0), R.mT.unsqueeze(0)).reshape(k, -1).mT * inv_sig.unsqueeze(-1)
            else:
                w = w[0]
            w, _, _ = _kernels.cap_solve(self._gram().reshape(-1, k, k), w.unsqueeze(-1))
            w = w.squeeze(-1)  # (B, k)
            x = _kernels.gemm3x(w.unsqueeze(0), U.unsqueeze(0), trans_b=True, row_alpha=(-inv_sig).unsqueeze(0),
                                E=R.unsqueeze(0), row_beta=inv_sig.unsqueeze(0))
            if x is None:
                S = _kernels.matmul_nn(U.unsqueeze(0), w.mT.unsqueeze(0))[0].mT  # (B, N) = (U w^T)^T
                x = (R - S) * inv_sig.unsqueeze(-1)
            else:
                x = x[0]
            return x.reshape(*self.batch_shape, n, 1)
        dinv_b = _kernels.scale_rows(rhs, d, "div")  # D^-1 b
        w = _kernels.tn_matmul(U, dinv_b)  # U^T D^-1 b
        w, _, _ = _kernels.cap_solve(self._gram(), w)  # (I + U^T D^-1 U)^-1 .
        res = _kernels.scale_rows(_kernels.matmul_nn(U, w), d, "div")  # D^-1 U .
        return dinv_b.sub_(res)

    def _logdet(self):  # :95-101
        U = self._linear_op._root_tensor()
        k = U.shape[-1]
        dummy = torch.zeros(*self.batch_shape, k, 1, dtype=self.dtype, device=self.device)
        _, logdet_cap, _ = _kernels.cap_solve(self._gram(), dummy)
        return logdet_cap + self._diag_tensor.logdet()

    def inv_quad_logdet(self, inv_quad_rhs=None, logdet=False, reduce_inv_quad=True):  # :114-160
        if not self.is_square:
            raise RuntimeError(
                "inv_quad_logdet only operates on (batches of) square (positive semi-definite) LinearOperators. "
                "Got a {} of size {}.".format(self.__class__.__name__, self.size())
            )
        if inv_quad_rhs is not None:
            if self.dim() == 2 and inv_quad_rhs.dim() == 1:
                if self.shape[-1] != inv_quad_rhs.numel():
                    raise RuntimeError(
                        "LinearOperator (size={}) cannot be multiplied with right-hand-side Tensor (size={}).".format(
                            self.shape, inv_quad_rhs.shape
                        )
                    )
            elif self.dim() != inv_quad_rhs.dim():
                raise RuntimeError(
                    "LinearOperator (size={}) and right-hand-side Tensor (size={}) should have the same number "
                    "of dimensions.".format(self.shape, inv_quad_rhs.shape)
                )
            elif self.batch_shape != inv_quad_rhs.shape[:-2] or self.shape[-1] != inv_quad_rhs.shape[-2]:
                raise RuntimeError(
                    "LinearOperator (size={}) cannot be multiplied with right-hand-side Tensor (size={}).".format(
                        self.shape, inv_quad_rhs.shape
                    )
                )
        inv_quad_term, logdet_term = None, None
        if inv_quad_rhs is not None:
            rhs = inv_quad_rhs.unsqueeze(-1) if inv_quad_rhs.dim() == 1 else inv_quad_rhs
            sol = self._solve(rhs)
            inv_quad_term = _kernels.col_dots(rhs, 0, sol, 0, rhs.shape[-1])
            if inv_quad_rhs.dim() == 1:
                inv_quad_term = inv_quad_term.squeeze(-1)
            elif reduce_inv_quad:
                inv_quad_term = inv_quad_term.sum(dim=-1)
        if logdet:
            logdet_term = self._logdet()
        return inv_quad_term, logdet_term

    def solve(self, right_tensor, left_tensor=None):  # :162-193
        if not self.is_square:
            raise RuntimeError(
                "solve only operates on (batches of) square (positive semi-definite) LinearOperators. "
                "Got a {} of size {}.".format(self.__class__.__name__, self.size())
            )
        if self.dim() == 2 and right_tensor.dim() == 1:
            if self.shape[-1] != right_tensor.numel():
                raise RuntimeError(
                    "LinearOperator (size={}) cannot be multiplied with right-hand-side Tensor (size={}).".format(
                        self.shape, right_tensor.shape
                    )
                )
        squeeze = right_tensor.ndimension() == 1
        rhs = right_tensor.unsqueeze(-1) if squeeze else right_tensor
        sol = self._solve(rhs)
        if squeeze:
            sol = sol.squeeze(-1)
        return sol if left_tensor is None else left_tensor @ sol

    def logdet(self):
        return self._logdet()

    def __add__(self, other):
        if isinstance(other, DiagLinearOperator):
            return self.__class__(self._linear_op, self._diag_tensor + other)
        return AddedDiagLinearOperator(self._linear_op + other, self._diag_tensor)


__all__ = ["LowRankRootAddedDiagLinearOperator"]

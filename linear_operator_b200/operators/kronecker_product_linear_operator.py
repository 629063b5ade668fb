"""KroneckerProductLinearOperator (reference: operators/kronecker_product_linear_operator.py)."""
from __future__ import annotations

import operator
from functools import reduce

import torch

from .. import _kernels
from ._linear_operator import LinearOperator
from .dense_linear_operator import to_linear_operator


class KroneckerProductLinearOperator(LinearOperator):
    """``K1 (x) K2 (x) ...`` with the first factor slowest in the row index.  ``_matmul`` is a chain of fused mode
    products (reference :34-45 does a bmm plus a transposing copy per factor)."""

    def __init__(self, *linear_ops):
        try:
            linear_ops = tuple(to_linear_operator(op) for op in linear_ops)
        except TypeError:
            raise RuntimeError("KroneckerProductLinearOperator is intended to wrap lazy tensors.")
        batch_shape = torch.broadcast_shapes(*[op.batch_shape for op in linear_ops])
        linear_ops = tuple(op if op.batch_shape == batch_shape else op._expand_batch(batch_shape) for op in linear_ops)
        super().__init__(*linear_ops)
        self.linear_ops = linear_ops

    def _factor_tensors(self):
        return [op.to_dense() for op in self.linear_ops]

    def _size(self):
        rows = reduce(operator.mul, [op.size(-2) for op in self.linear_ops], 1)
        cols = reduce(operator.mul, [op.size(-1) for op in self.linear_ops], 1)
        return torch.Size((*self.linear_ops[0].batch_shape, rows, cols))

    def _matmul(self, rhs):  # :272-284
        squeeze = rhs.dim() == 1
        if squeeze:
            rhs = rhs.unsqueeze(-1)
        res = _kernels.kron_matmul(self._factor_tensors(), rhs)
        return res.squeeze(-1) if squeeze else res

    def add_diagonal(self, diag):  # :116-150
        from .diag_linear_operator import ConstantDiagLinearOperator, DiagLinearOperator
        from .kronecker_product_added_diag_linear_operator import KroneckerProductAddedDiagLinearOperator

        if not self.is_square:
            raise RuntimeError("add_diag only defined for square matrices")
        if diag.dim() == 0:
            diag_op = ConstantDiagLinearOperator(diag.unsqueeze(-1), diag_shape=self.shape[-1])
        elif diag.shape[-1] == 1:
            diag_op = ConstantDiagLinearOperator(diag, diag_shape=self.shape[-1])
        else:
            try:
                expanded = diag.expand(self.shape[:-1])
            except RuntimeError:
                raise RuntimeError(
                    "add_diagonal for LinearOperator of size {} received invalid diagonal of size {}.".format(
                        self.shape, diag.shape
                    )
                ) from None
            diag_op = DiagLinearOperator(expanded)
        return KroneckerProductAddedDiagLinearOperator(self, diag_op)

    def __add__(self, other):  # :98-114
        from .diag_linear_operator import ConstantDiagLinearOperator, DiagLinearOperator
        from .kronecker_product_added_diag_linear_operator import KroneckerProductAddedDiagLinearOperator

        if isinstance(other, ConstantDiagLinearOperator):
            return KroneckerProductAddedDiagLinearOperator(self, other)
        if isinstance(other, DiagLinearOperator):
            return self.add_diagonal(other._diagonal())
        return super().__add__(other)

    def _matmul_add_diag(self, rhs, diag, want_dots=False):
        """K X + d (.) X with the diagonal folded into the chain's last pass (AddedDiagLinearOperator._matmul);
        ``want_dots``: and linear_cg's partial <X, Y> sums out of the same pass."""
        return _kernels.kron_matmul(self._factor_tensors(), rhs, d=diag, want_dots=want_dots)

    def _matmul_closure(self):
        def closure(v):
            return self._matmul(v)

        closure.fused = lambda v: self._matmul_add_diag(v, None, want_dots=True)
        return closure

    def _transpose_nonbatch(self):
        return self.__class__(*(op._transpose_nonbatch() for op in self.linear_ops))

    def _diagonal(self):  # :20-27, :192-196
        d = self.linear_ops[0]._diagonal()
        for op in self.linear_ops[1:]:
            dn = op._diagonal()
            d = (d.unsqueeze(-1) * dn.unsqueeze(-2)).reshape(*d.shape[:-1], -1)
        return d

    def _expand_batch(self, batch_shape):
        return self.__class__(*[op._expand_batch(batch_shape) for op in self.linear_ops])

    def _get_indices(self, row_index, col_index, *batch_indices):  # :198-216
        row_factor = self.size(-2)
        col_factor = self.size(-1)
        res = None
        for op in self.linear_ops:
            row_factor //= op.size(-2)
            col_factor //= op.size(-1)
            sub = op._get_indices(
                torch.div(row_index, row_factor, rounding_mode="floor").fmod(op.size(-2)),
                torch.div(col_index, col_factor, rounding_mode="floor").fmod(op.size(-1)),
                *batch_indices,
            )
            res = sub if res is None else (sub * res)
        return res

    def _pivoted_cholesky(self, rank, error_tol):
        return _kernels.pivoted_cholesky_kron(self._factor_tensors(), self.batch_shape, rank, error_tol)


__all__ = ["KroneckerProductLinearOperator"]

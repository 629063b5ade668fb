"""SumLinearOperator / PsdSumLinearOperator (reference: operators/sum_linear_operator.py, psd_sum_linear_operator.py)."""
from __future__ import annotations

import torch

from ._linear_operator import LinearOperator
from .dense_linear_operator import to_linear_operator


class SumLinearOperator(LinearOperator):
    def __init__(self, *linear_ops, **kwargs):
        linear_ops = [to_linear_operator(op) for op in linear_ops]
        shape = torch.broadcast_shapes(*[op.batch_shape for op in linear_ops])
        linear_ops = [op if op.batch_shape == shape else op._expand_batch(shape) for op in linear_ops]
        super().__init__(*linear_ops, **kwargs)
        self.linear_ops = tuple(linear_ops)

    def _matmul(self, rhs):  # :47-51
        # each term is one structured kernel; the running sum is accumulated term by term
        out = self.linear_ops[0]._matmul(rhs)
        for op in self.linear_ops[1:]:
            out = out.add_(op._matmul(rhs))
        return out

    def _bilinear_derivative(self, left_vecs, right_vecs):  # :59-62
        return tuple(
            var for linear_op in self.linear_ops for var in linear_op._bilinear_derivative(left_vecs, right_vecs)
        )

    def _size(self):
        return self.linear_ops[0].size()

    def _transpose_nonbatch(self):
        return self.__class__(*[op._transpose_nonbatch() for op in self.linear_ops])

    def _diagonal(self):  # :24-25
        d = self.linear_ops[0]._diagonal().clone()
        for op in self.linear_ops[1:]:
            d = d + op._diagonal()
        return d

    def _expand_batch(self, batch_shape):
        return self.__class__(*[op._expand_batch(batch_shape) for op in self.linear_ops])

    def _get_indices(self, row_index, col_index, *batch_indices):
        res = self.linear_ops[0]._get_indices(row_index, col_index, *batch_indices)
        for op in self.linear_ops[1:]:
            res = res + op._get_indices(row_index, col_index, *batch_indices)
        return res

    def to_dense(self):
        return sum(op.to_dense() for op in self.linear_ops)


class PsdSumLinearOperator(SumLinearOperator):
    """A sum of PSD operators; samples are sums of the terms' samples, drawn in operand order
    (reference psd_sum_linear_operator.py:15-18)."""

    def zero_mean_mvn_samples(self, num_samples):
        return sum(op.zero_mean_mvn_samples(num_samples) for op in self.linear_ops)


__all__ = ["SumLinearOperator", "PsdSumLinearOperator"]

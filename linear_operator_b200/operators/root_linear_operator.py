"""RootLinearOperator / LowRankRootLinearOperator (reference: operators/root_linear_operator.py,
low_rank_root_linear_operator.py)."""
from __future__ import annotations

import torch

from .. import _kernels
from ._linear_operator import LinearOperator
from .dense_linear_operator import to_linear_operator


class RootLinearOperator(LinearOperator):
    """``R R^T`` for a root ``R`` of shape ``(*batch, N, r)``."""

    def __init__(self, root):
        root = to_linear_operator(root)
        super().__init__(root)
        self.root = root

    def _root_tensor(self):
        return self.root.to_dense()

    def _matmul(self, rhs):  # :68-72: R (R^T rhs)
        R = self._root_tensor()
        return _kernels.matmul_nn(R, _kernels.tn_matmul(R, rhs))

    def _bilinear_derivative(self, left_vecs, right_vecs):
        """d/dR sum_i u_i^T R R^T v_i = U (R^T V)^T + V (R^T U)^T -- what the reference's autograd default (:336-393)
        yields for ``_matmul = R (R^T .)``; the root is itself an operator and receives both terms through ITS hook."""
        if left_vecs.dim() == 1:
            left_vecs, right_vecs = left_vecs.unsqueeze(-1), right_vecs.unsqueeze(-1)
        R = self._root_tensor()
        rt_right = _kernels.tn_matmul(R, right_vecs)  # (*, r, C)
        rt_left = _kernels.tn_matmul(R, left_vecs)
        a = self.root._bilinear_derivative(left_vecs, rt_right)
        b = self.root._bilinear_derivative(right_vecs, rt_left)
        return tuple(None if x is None else x + y for x, y in zip(a, b))

    def _size(self):
        n = self.root.size(-2)
        return torch.Size((*self.root.batch_shape, n, n))

    def _transpose_nonbatch(self):
        return self

    def _diagonal(self):  # :31-36: row norms of the root
        R = self._root_tensor()
        return (R * R).sum(-1)

    def _expand_batch(self, batch_shape):
        return self.__class__(self.root._expand_batch(batch_shape))

    def _get_indices(self, row_index, col_index, *batch_indices):  # :47-58
        R = self._root_tensor()
        left = R[(*batch_indices, row_index)]
        right = R[(*batch_indices, col_index)]
        return (left * right).sum(-1)

    def to_dense(self):
        R = self._root_tensor()
        eye = torch.eye(R.size(-2), dtype=R.dtype, device=R.device).expand(*R.shape[:-2], R.size(-2), R.size(-2))
        return self._matmul(eye.contiguous())

    def zero_mean_mvn_samples(self, num_samples):
        """R eps with eps = randn(*batch, r, S), returned as (S, *batch, N) (operators/_linear_operator.py:2779-2791)."""
        R = self._root_tensor()
        base = torch.randn(*self.batch_shape, R.size(-1), num_samples, dtype=self.dtype, device=self.device)
        samples = _kernels.matmul_nn(R, base)  # (*batch, N, S)
        return samples.permute(-1, *range(self.dim() - 1)).contiguous()


class LowRankRootLinearOperator(RootLinearOperator):
    """A RootLinearOperator whose root has few columns; adding a diagonal yields the Woodbury operator
    (reference low_rank_root_linear_operator.py:52-64)."""

    def add_diagonal(self, diag):
        from .diag_linear_operator import ConstantDiagLinearOperator, DiagLinearOperator
        from .low_rank_root_added_diag_linear_operator import LowRankRootAddedDiagLinearOperator

        if not self.is_square:
            raise RuntimeError("add_diag only defined for square matrices")
        n = self.size(-1)
        if diag.dim() == 0:
            diag_op = ConstantDiagLinearOperator(diag.unsqueeze(-1), diag_shape=n)
        elif diag.shape[-1] == 1:
            diag_op = ConstantDiagLinearOperator(diag, diag_shape=n)
        else:
            diag_op = DiagLinearOperator(diag.expand(*self.batch_shape, n))
        return LowRankRootAddedDiagLinearOperator(self, diag_op)

    def __add__(self, other):
        from .diag_linear_operator import DiagLinearOperator
        from .low_rank_root_added_diag_linear_operator import LowRankRootAddedDiagLinearOperator

        if isinstance(other, DiagLinearOperator):
            return LowRankRootAddedDiagLinearOperator(self, other)
        return super().__add__(other)


__all__ = ["RootLinearOperator", "LowRankRootLinearOperator"]

"""DiagLinearOperator / ConstantDiagLinearOperator (reference: operators/diag_linear_operator.py)."""
from __future__ import annotations

import torch

from .. import _kernels
from ._linear_operator import LinearOperator


class DiagLinearOperator(LinearOperator):
    """Diagonal operator defined by ``diag`` of shape ``(*batch, N)`` (reference :23-39)."""

    def __init__(self, diag):
        super().__init__(diag)
        self._diag = diag

    def _size(self):  # :195-196
        return self._diag.shape + self._diag.shape[-1:]

    def _transpose_nonbatch(self):
        return self

    def _diagonal(self):  # :153-154
        return self._diag

    def _matmul(self, rhs):  # :203-206  d (.) rhs
        squeeze = rhs.dim() == 1
        if squeeze:
            rhs = rhs.unsqueeze(-1)
        out = _kernels.scale_rows(rhs, self._diag, "mul")
        return out.squeeze(-1) if squeeze else out

    def _expand_batch(self, batch_shape):
        return self.__class__(self._diag.expand(*batch_shape, self._diag.size(-1)))

    def _bilinear_derivative(self, left_vecs, right_vecs):  # :37-45
        if not self._diag.requires_grad:
            return (None,)
        if left_vecs.dim() == 1:
            left_vecs, right_vecs = left_vecs.unsqueeze(-1), right_vecs.unsqueeze(-1)
        res = _kernels.bilinear_diag(left_vecs, right_vecs)
        if res.shape != self._diag.shape:
            res = res.sum_to_size(self._diag.shape)
        return (res,)

    def _get_indices(self, row_index, col_index, *batch_indices):  # :73-78
        res = self._diag[(*batch_indices, row_index)]
        return res * torch.eq(row_index, col_index).to(device=res.device, dtype=res.dtype)

    def add_diagonal(self, added_diag):
        shape = torch.broadcast_shapes(self._diag.shape, added_diag.shape)
        return DiagLinearOperator(self._diag.expand(shape) + added_diag.expand(shape))

    def __add__(self, other):  # :50-58
        if isinstance(other, DiagLinearOperator):
            return self.add_diagonal(other._diag)
        from .added_diag_linear_operator import AddedDiagLinearOperator

        return AddedDiagLinearOperator(other, self) if isinstance(other, LinearOperator) else super().__add__(other)

    def inverse(self):  # :220-222
        return self.__class__(self._diag.reciprocal())

    def logdet(self):  # :232-233
        return self._diag.log().sum(-1)

    def to_dense(self):
        return torch.diag_embed(self._diag)

    def solve(self, right_tensor, left_tensor=None):  # :250-262
        squeeze = right_tensor.dim() == 1
        rhs = right_tensor.unsqueeze(-1) if squeeze else right_tensor
        res = _kernels.scale_rows(rhs, self._diag, "div")
        res = res.squeeze(-1) if squeeze else res
        return res if left_tensor is None else left_tensor @ res

    def zero_mean_mvn_samples(self, num_samples):  # :273-277: randn(S, *batch, N) * sqrt(d)
        base = torch.randn(num_samples, *self._diag.shape, dtype=self.dtype, device=self.device)
        return base * self._diag.sqrt()


class ConstantDiagLinearOperator(DiagLinearOperator):
    """Diagonal with one value per batch element: ``diag_values`` is ``(*batch, 1)`` (reference :300-350)."""

    def __init__(self, diag_values, diag_shape):
        LinearOperator.__init__(self, diag_values, diag_shape=diag_shape)
        self.diag_values = diag_values
        self.diag_shape = diag_shape

    def _check_args(self, diag_values, diag_shape):
        if not torch.is_tensor(diag_values):
            return f"diag_values must be a Tensor, got {type(diag_values)}"
        if diag_values.dim() < 1 or diag_values.size(-1) != 1:
            return f"diag_values must have a trailing dimension of size 1, got {tuple(diag_values.shape)}"

    @property
    def _diag(self):  # :346-350: a stride-0 expanded view
        return self.diag_values.expand(*self.diag_values.shape[:-1], self.diag_shape)

    def _expand_batch(self, batch_shape):
        return self.__class__(self.diag_values.expand(*batch_shape, 1), diag_shape=self.diag_shape)

    def _bilinear_derivative(self, left_vecs, right_vecs):  # :337-344
        if not self.diag_values.requires_grad:
            return (None,)
        if left_vecs.dim() == 1:
            left_vecs, right_vecs = left_vecs.unsqueeze(-1), right_vecs.unsqueeze(-1)
        res = _kernels.bilinear_diag(left_vecs, right_vecs).sum(-1, keepdim=True)
        if res.shape != self.diag_values.shape:
            res = res.sum_to_size(self.diag_values.shape)
        return (res,)

    def add_diagonal(self, added_diag):
        if added_diag.dim() == 0 or added_diag.size(-1) == 1:
            v = added_diag.reshape(*added_diag.shape[:-1], 1) if added_diag.dim() else added_diag.reshape(1)
            shape = torch.broadcast_shapes(self.diag_values.shape, v.shape)
            return ConstantDiagLinearOperator(self.diag_values.expand(shape) + v.expand(shape), self.diag_shape)
        return DiagLinearOperator(self._diag + added_diag)

    def __add__(self, other):
        if isinstance(other, ConstantDiagLinearOperator):
            if other.shape[-1] != self.shape[-1]:
                raise RuntimeError(f"Trailing batch shapes must match for adding two ConstantDiagLinearOperators.")
            return ConstantDiagLinearOperator(self.diag_values + other.diag_values, self.diag_shape)
        return super().__add__(other)

    def inverse(self):
        return ConstantDiagLinearOperator(self.diag_values.reciprocal(), self.diag_shape)

    def logdet(self):
        return self.diag_values.squeeze(-1).log() * self.diag_shape


__all__ = ["DiagLinearOperator", "ConstantDiagLinearOperator"]

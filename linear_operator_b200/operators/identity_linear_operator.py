"""IdentityLinearOperator (reference: operators/identity_linear_operator.py) -- the ``precond_lt`` of operators without
a preconditioner (operators/_linear_operator.py:1774-1783)."""
from __future__ import annotations

import torch

from ._linear_operator import LinearOperator
from .diag_linear_operator import ConstantDiagLinearOperator


class IdentityLinearOperator(ConstantDiagLinearOperator):
    def __init__(self, diag_shape, batch_shape=torch.Size([]), dtype=None, device=None):
        one = torch.tensor(1.0, dtype=dtype, device=device)
        LinearOperator.__init__(self, diag_shape=diag_shape, batch_shape=batch_shape, dtype=dtype, device=device)
        self.diag_values = one.expand(*batch_shape, 1)
        self.diag_shape = diag_shape
        self._batch_shape = torch.Size(batch_shape)
        self._dtype = dtype
        self._device = device

    def _check_args(self, *args, **kwargs):
        return None

    @property
    def batch_shape(self):
        return self._batch_shape

    @property
    def dtype(self):
        return self._dtype

    @property
    def device(self):
        return self._device

    def _size(self):
        return torch.Size([*self._batch_shape, self.diag_shape, self.diag_shape])

    def _matmul(self, rhs):
        return rhs.expand(*torch.broadcast_shapes(self._batch_shape, rhs.shape[:-2]), *rhs.shape[-2:]).clone()

    def _expand_batch(self, batch_shape):
        return IdentityLinearOperator(self.diag_shape, batch_shape, self._dtype, self._device)

    def zero_mean_mvn_samples(self, num_samples):  # :262-266
        return torch.randn(num_samples, *self._batch_shape, self.diag_shape, dtype=self._dtype, device=self._device)

    def _bilinear_derivative(self, left_vecs, right_vecs):
        return ()

    def logdet(self):
        return torch.zeros(self._batch_shape, dtype=self._dtype, device=self._device)

    def solve(self, right_tensor, left_tensor=None):
        return right_tensor if left_tensor is None else left_tensor @ right_tensor


__all__ = ["IdentityLinearOperator"]

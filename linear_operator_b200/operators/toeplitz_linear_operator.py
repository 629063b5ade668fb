"""ToeplitzLinearOperator (reference: operators/toeplitz_linear_operator.py, utils/toeplitz.py)."""
from __future__ import annotations

import torch

from .. import _kernels
from ._linear_operator import LinearOperator


class ToeplitzLinearOperator(LinearOperator):
    """Symmetric Toeplitz operator defined by its first column ``(*batch, N)``.  ``_matmul`` embeds it in a power-of-two
    real circulant and uses cuFFT R2C/C2R with custom pad / multiply / unpad kernels (reference utils/toeplitz.py
    :131-149 uses an odd-length complex transform with two transposing copies)."""

    def __init__(self, column):
        super().__init__(column)
        self.column = column
        self._fc_cache = None

    def _spectrum(self):
        if self._fc_cache is None:
            self._fc_cache = _kernels.toeplitz_embed_fft(self.column)
        return self._fc_cache

    def _matmul(self, rhs):  # :42-46
        squeeze = rhs.dim() == 1
        if squeeze:
            rhs = rhs.unsqueeze(-1)
        res = _kernels.toeplitz_matmul(self.column, rhs, fc_cache=self._spectrum())
        return res.squeeze(-1) if squeeze else res

    def _matmul_add_diag(self, rhs, diag, want_dots=False):
        """T X + d (.) X with the diagonal folded into the un-padding kernel (``want_dots``: and linear_cg's partial
        <X, Y> sums out of the same pass)."""
        return _kernels.toeplitz_matmul(self.column, rhs, d=diag, fc_cache=self._spectrum(), want_dots=want_dots)

    def _matmul_closure(self):
        def closure(v):
            return self._matmul(v)

        closure.fused = lambda v: self._matmul_add_diag(v, None, want_dots=True)
        return closure

    def _bilinear_derivative(self, left_vecs, right_vecs):  # :55-66 -> utils/toeplitz.py:164-204
        if left_vecs.dim() == 1:
            left_vecs, right_vecs = left_vecs.unsqueeze(-1), right_vecs.unsqueeze(-1)
        res = _kernels.toeplitz_bilinear_derivative(left_vecs, right_vecs)
        if res.shape != self.column.shape:  # collapse expanded broadcast dimensions (:63-64)
            res = res.sum_to_size(self.column.shape)
        return (res,)

    def _size(self):
        return torch.Size((*self.column.shape, self.column.size(-1)))

    def _transpose_nonbatch(self):
        return ToeplitzLinearOperator(self.column)

    def _diagonal(self):  # :25-31
        return self.column[..., 0].unsqueeze(-1).expand(*self.column.shape)

    def _expand_batch(self, batch_shape):
        return self.__class__(self.column.expand(*batch_shape, self.column.size(-1)))

    def _get_indices(self, row_index, col_index, *batch_indices):  # :38-40
        return self.column[(*batch_indices, (row_index - col_index).abs())]

    def add_jitter(self, jitter_val=1e-3):  # :76-81
        jitter = torch.zeros_like(self.column)
        jitter.narrow(-1, 0, 1).fill_(jitter_val)
        return ToeplitzLinearOperator(self.column.add(jitter))

    def _pivoted_cholesky(self, rank, error_tol):
        return _kernels.pivoted_cholesky_toeplitz(self.column, rank, error_tol)


__all__ = ["ToeplitzLinearOperator"]

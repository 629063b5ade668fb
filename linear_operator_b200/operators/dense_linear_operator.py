"""DenseLinearOperator (reference: operators/dense_linear_operator.py)."""
from __future__ import annotations

import torch

from .. import _kernels
from ._linear_operator import LinearOperator


class DenseLinearOperator(LinearOperator):
    """Wraps an explicit ``(*batch, M, N)`` tensor.  ``_matmul`` is the dense kernel of ``liblob_b200``
    (reference :60-64 calls torch.matmul -> cuBLAS)."""

    def _check_args(self, tsr):
        if not torch.is_tensor(tsr):
            return "DenseLinearOperator must take a torch.Tensor; got {}".format(tsr.__class__.__name__)
        if tsr.dim() < 2:
            return "DenseLinearOperator expects a matrix (or batches of matrices) - got a Tensor of size {}.".format(
                tsr.shape
            )

    def __init__(self, tsr):
        super().__init__(tsr)
        self.tensor = tsr

    def _matmul(self, rhs):
        return _kernels.dense_matmul(self.tensor, rhs)

    def _matmul_closure(self):
        tsr = self.tensor

        def closure(v):
            return _kernels.dense_matmul(tsr, v)

        closure.fused = lambda v: _kernels.dense_matmul(tsr, v, want_dots=True)
        closure.graph_spec = (tsr, None)  # lets linear_cg replay small solves as one CUDA graph (settings.cuda_graphs)
        return closure

    def _bilinear_derivative(self, left_vecs, right_vecs):  # :69-71: left right^T, one rank-C outer-product kernel
        if not self.tensor.requires_grad:  # a (*batch, M, N) gradient nobody asked for is the costliest thing to skip
            return (None,)
        if left_vecs.dim() == 1:
            left_vecs, right_vecs = left_vecs.unsqueeze(-1), right_vecs.unsqueeze(-1)
        res = _kernels.bilinear_dense(left_vecs, right_vecs)
        if res.shape != self.tensor.shape:  # operator batch smaller than the vectors' batch: collapse the broadcast
            res = res.sum_to_size(self.tensor.shape)
        return (res,)

    def _size(self):
        return self.tensor.size()

    def _transpose_nonbatch(self):
        return DenseLinearOperator(self.tensor.mT)

    def _diagonal(self):  # :37-40
        return self.tensor.diagonal(dim1=-1, dim2=-2)

    def _expand_batch(self, batch_shape):  # :42-45
        return self.__class__(self.tensor.expand(*batch_shape, *self.matrix_shape))

    def _get_indices(self, row_index, col_index, *batch_indices):  # :47-50
        return self.tensor[(*batch_indices, row_index, col_index)]

    def _getitem(self, index):
        return self.__class__(self.tensor[index]) if self.tensor[index].dim() >= 2 else self.tensor[index]

    def to_dense(self):
        return self.tensor

    def _pivoted_cholesky(self, rank, error_tol):
        return _kernels.pivoted_cholesky_dense(self.tensor, rank, error_tol)


def to_linear_operator(obj):
    """Tensor -> DenseLinearOperator, LinearOperator -> itself (reference :107-120)."""
    if torch.is_tensor(obj):
        return DenseLinearOperator(obj)
    if isinstance(obj, LinearOperator):
        return obj
    raise TypeError("object of class {} cannot be made into a LinearOperator".format(obj.__class__.__name__))


__all__ = ["DenseLinearOperator", "to_linear_operator"]

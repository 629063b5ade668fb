"""Diagonalization Function: partial Lanczos diagonalisation Q S Q^T ~ A (reference: functions/_diagonalization.py:10-95)."""
from __future__ import annotations

import torch
from torch.autograd import Function

from .. import _kernels
from ._root_decomposition import _lanczos_eig


class Diagonalization(Function):
    @staticmethod
    def forward(ctx, representation_tree, device, dtype, matrix_shape, max_iter, batch_shape, *matrix_args):
        linear_op = representation_tree(*matrix_args)
        q_mat, eigenvalues, eigenvectors = _lanczos_eig(ctx, linear_op, max_iter, dtype, device, matrix_shape,
                                                        torch.Size(batch_shape), None)
        q_mat = _kernels.matmul_nn(q_mat, eigenvectors).squeeze(0)  # :55,:62
        eigenvalues = eigenvalues.squeeze(0)
        ctx.save_for_backward(*matrix_args, q_mat, eigenvalues)
        return eigenvalues, q_mat

    @staticmethod
    def backward(ctx, evals_grad_output, evecs_grad_output):
        """Reference :69-95 (Ionescu et al. 2015): dL/dM = Q (K~^T o (Q^T dL/dQ)) Q^T + Q diag(dL/dS) Q^T, returned for
        the first (dense) matrix argument as the reference does; both terms are rank-k outer-product kernels."""
        q_mat, eigenvalues = ctx.saved_tensors[-2:]
        kmat = (eigenvalues.unsqueeze(-1) - eigenvalues.unsqueeze(-2) + 1e-10).reciprocal()
        torch.diagonal(kmat, dim1=-1, dim2=-2).zero_()
        out = None
        if evecs_grad_output is not None:
            inner = kmat.mT * _kernels.tn_matmul(q_mat, evecs_grad_output)  # (k x k)
            out = _kernels.bilinear_dense(_kernels.matmul_nn(q_mat, inner), q_mat)
        if evals_grad_output is not None:
            out = _kernels.bilinear_dense(q_mat, q_mat, w=evals_grad_output, out=None if out is None else
                                          out.contiguous())
        return tuple([None] * 6 + [out] + [None] * (len(ctx.saved_tensors) - 3))

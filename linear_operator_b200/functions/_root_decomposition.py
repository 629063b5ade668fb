"""RootDecomposition Function: Lanczos root / inverse root R R^T ~ A, R_inv R_inv^T ~ A^-1
(reference: functions/_root_decomposition.py:12-180).  Lanczos with full re-orthogonalisation is one fused kernel per
iteration (csrc/lanczos.cu), the tridiagonal eigendecomposition runs on the device (csrc/tridiag.cu; the reference ships
it to CPU LAPACK), Q V diag(sqrt(lambda)) is one skinny matmul."""
from __future__ import annotations

import torch
from torch.autograd import Function

from .. import _kernels, settings
from ..utils import lanczos


def _lanczos_eig(ctx_like, linear_op, max_iter, dtype, device, matrix_shape, batch_shape, initial_vectors):
    """Lanczos -> jittered tridiagonal eigendecomposition; returns q_mat, eigenvalues, eigenvectors with a leading probe
    dimension and (at least) the operator's batch shape (:47-75)."""
    q_mat, t_mat = lanczos.lanczos_tridiag(
        linear_op._matmul, max_iter, dtype=dtype, device=device, matrix_shape=matrix_shape, batch_shape=batch_shape,
        init_vecs=initial_vectors,
    )
    if t_mat.ndimension() == 2 + len(batch_shape):  # one probe vector: the probe dimension was squeezed
        q_mat = q_mat.unsqueeze(0)
        t_mat = t_mat.unsqueeze(0)
    mins = torch.diagonal(t_mat, dim1=-1, dim2=-2).min(dim=-1, keepdim=True)[0].unsqueeze(-1)
    jitter_mat = (settings.tridiagonal_jitter.value() * mins) * torch.eye(
        t_mat.size(-1), device=t_mat.device, dtype=t_mat.dtype
    ).expand_as(t_mat)
    eigenvalues, eigenvectors = lanczos.lanczos_tridiag_to_diag(t_mat + jitter_mat)
    return q_mat, eigenvalues, eigenvectors


class RootDecomposition(Function):
    @staticmethod
    def forward(ctx, representation_tree, max_iter, dtype, device, batch_shape, matrix_shape, root, inverse,
                initial_vectors, *matrix_args):
        ctx.representation_tree = representation_tree
        ctx.inverse = inverse
        linear_op = representation_tree(*matrix_args)
        q_mat, eigenvalues, eigenvectors = _lanczos_eig(ctx, linear_op, max_iter, dtype, device, matrix_shape,
                                                        torch.Size(batch_shape), initial_vectors)
        n_probes = q_mat.size(0)
        root_evals = eigenvalues.sqrt()
        # Q V diag(s) = Q (V diag(s)): scale the k x k eigenvector matrix, then one (N x k)(k x k) product (:78-88)
        empty = torch.empty(0, dtype=q_mat.dtype, device=q_mat.device)
        root_t, inverse_t = empty, empty
        if inverse:
            inverse_t = _kernels.matmul_nn(q_mat, eigenvectors / root_evals.unsqueeze(-2))
        if root:
            root_t = _kernels.matmul_nn(q_mat, eigenvectors * root_evals.unsqueeze(-2))
        q_rot = _kernels.matmul_nn(q_mat, eigenvectors) if any(ctx.needs_input_grad) else empty
        if n_probes == 1:  # :101-105
            root_t = root_t.squeeze(0) if root_t.numel() else root_t
            inverse_t = inverse_t.squeeze(0) if inverse_t.numel() else inverse_t
            q_rot = q_rot.squeeze(0) if q_rot.numel() else q_rot
            root_evals = root_evals.squeeze(0)
        ctx.save_for_backward(*matrix_args, q_rot, root_evals, inverse_t)
        return root_t, inverse_t

    @staticmethod
    def backward(ctx, root_grad_output, inverse_grad_output):
        """Reference :107-180 (Murray, "Differentiation of the Cholesky decomposition", applied to the Lanczos root):
        the operator gradient is _bilinear_derivative(left_factor, R^-T / 2)."""
        if not any(ctx.needs_input_grad):
            return tuple([None] * len(ctx.needs_input_grad))

        def is_empty(t):
            return t is None or t.numel() == 0 or (t.numel() == 1 and t.reshape(-1)[0] == 0)

        root_grad_output = None if is_empty(root_grad_output) else root_grad_output
        inverse_grad_output = None if is_empty(inverse_grad_output) else inverse_grad_output
        *matrix_args, q_mat, root_evals, inverse = ctx.saved_tensors
        is_batch = False
        if root_grad_output is not None:
            if root_grad_output.ndimension() == 2 and q_mat.ndimension() > 2:
                root_grad_output, is_batch = root_grad_output.unsqueeze(0), True
            if root_grad_output.ndimension() == 3 and q_mat.ndimension() > 3:
                root_grad_output, is_batch = root_grad_output.unsqueeze(0), True
        if inverse_grad_output is not None:
            if inverse_grad_output.ndimension() == 2 and q_mat.ndimension() > 2:
                inverse_grad_output, is_batch = inverse_grad_output.unsqueeze(0), True
            if inverse_grad_output.ndimension() == 3 and q_mat.ndimension() > 3:
                inverse_grad_output, is_batch = inverse_grad_output.unsqueeze(0), True
        linear_op = ctx.representation_tree(*matrix_args)
        if not ctx.inverse:
            inverse = q_mat / root_evals.unsqueeze(-2)
        left_factor = torch.zeros_like(inverse)
        if root_grad_output is not None:
            left_factor.add_(root_grad_output)
        if inverse_grad_output is not None:  # - R^-T G^T R^-T  (:152-154)
            inner = _kernels.tn_matmul(inverse_grad_output, inverse)  # G^T R_inv: (k x k)
            left_factor.sub_(_kernels.matmul_nn(inverse, inner))
        right_factor = inverse / 2.0
        if is_batch:  # probe dimension folded into the columns (:160-164)
            left_factor = left_factor.permute(1, 0, 2, 3).contiguous().view(inverse.size(1), -1, left_factor.size(-1))
            right_factor = right_factor.permute(1, 0, 2, 3).contiguous().view(inverse.size(1), -1, right_factor.size(-1))
        res = linear_op._bilinear_derivative(left_factor.contiguous(), right_factor.contiguous())
        return tuple([None] * 9 + list(res))

"""PivotedCholesky Function (reference: functions/_pivoted_cholesky.py:13-105).  The batched pivoted Cholesky
runs in ``csrc/pivchol.cu``.  Dense / Kronecker / Toeplitz operators hand the kernels a device functor as row source, so
the reference's per-step ``__getitem__`` gathers and host synchronisations (:57-98) disappear; every other operator
(Root, Sum, user classes) takes the reference's generic route -- one ``_get_indices`` row gather per step on
device-resident pivot indices -- through the same kernels (``_kernels.pivoted_cholesky_rows``)."""
from __future__ import annotations

import torch
from torch.autograd import Function

from .. import settings

_BACKWARD_MSG = (
    "The backward pass of {} (reference functions/{}) is a 'next' row of the hot-path scope table "
    "(SURVEY.md section 8f) and is not built yet; call it on tensors that do not require grad or under torch.no_grad()."
)


class PivotedCholesky(Function):
    @staticmethod
    def forward(ctx, representation_tree, max_iter, error_tol, *matrix_args):
        matrix = representation_tree(*matrix_args)
        if error_tol is None:
            error_tol = settings.preconditioner_tolerance.value()
        if settings.verbose_linalg.on():
            settings.verbose_linalg.logger.debug(
                f"Running Pivoted Cholesky on a {matrix.shape} RHS for {max_iter} iterations."
            )
        L, perm = matrix._pivoted_cholesky(max_iter, error_tol)
        ctx.mark_non_differentiable(perm)
        return L, perm

    @staticmethod
    def backward(ctx, grad_output, _):
        raise NotImplementedError(_BACKWARD_MSG.format("PivotedCholesky", "_pivoted_cholesky.py:107-147"))

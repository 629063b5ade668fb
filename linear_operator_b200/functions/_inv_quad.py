"""InvQuad Function (reference: functions/_inv_quad.py:10-61)."""
from __future__ import annotations

import torch
from torch.autograd import Function

from .. import _kernels, settings
from ._pivoted_cholesky import _BACKWARD_MSG


def _solve(linear_op, rhs):
    if (
        settings.fast_computations.solves.off()
        or settings.fast_computations.log_prob.off()
        or linear_op.size(-1) <= settings.max_cholesky_size.value()
    ):
        return torch.cholesky_solve(rhs, linear_op.cholesky())
    with torch.no_grad():
        preconditioner = linear_op._solve_preconditioner()
    return linear_op._solve(rhs, preconditioner)


class InvQuad(Function):
    @staticmethod
    def forward(ctx, representation_tree, *args):
        inv_quad_rhs, *matrix_args = args
        linear_op = representation_tree(*matrix_args)
        if inv_quad_rhs.ndimension() == 1:
            inv_quad_rhs = inv_quad_rhs.unsqueeze(-1)
        solves = _solve(linear_op, inv_quad_rhs)
        if solves.is_cuda:
            term = _kernels.col_dots(solves, 0, inv_quad_rhs.expand_as(solves), 0, solves.shape[-1])
        else:
            term = (solves * inv_quad_rhs).sum(-2)
        ctx.mark_non_differentiable(term)
        return term

    @staticmethod
    def backward(ctx, grad_output):
        raise NotImplementedError(_BACKWARD_MSG.format("InvQuad", "_inv_quad.py:63-93"))

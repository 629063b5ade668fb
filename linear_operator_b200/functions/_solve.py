"""Solve Function (reference: functions/_solve.py:10-68)."""
from __future__ import annotations

import torch
from torch.autograd import Function

from .. import settings
from ._pivoted_cholesky import _BACKWARD_MSG


def _solve(linear_op, rhs):
    """Dense Cholesky below ``max_cholesky_size`` (not the Krylov path), preconditioned CG above (:10-22)."""
    if settings.fast_computations.solves.off() or linear_op.size(-1) <= settings.max_cholesky_size.value():
        return torch.cholesky_solve(rhs, linear_op.cholesky())
    with torch.no_grad():
        preconditioner = linear_op._solve_preconditioner()
    return linear_op._solve(rhs, preconditioner)


class Solve(Function):
    @staticmethod
    def forward(ctx, representation_tree, has_left, *args):
        if has_left:
            left_tensor, right_tensor, *matrix_args = args
        else:
            left_tensor = None
            right_tensor, *matrix_args = args
        linear_op = representation_tree(*matrix_args)
        is_vector = right_tensor.ndimension() == 1
        if is_vector:
            right_tensor = right_tensor.unsqueeze(-1)
        if has_left:  # :48-52
            rhs = torch.cat([left_tensor.mT, right_tensor], -1)
            solves = _solve(linear_op, rhs)
            res = left_tensor @ solves[..., left_tensor.size(-2):]
        else:
            res = _solve(linear_op, right_tensor)
        if is_vector:
            res = res.squeeze(-1)
        ctx.mark_non_differentiable(res)
        return res

    @staticmethod
    def backward(ctx, grad_output):
        raise NotImplementedError(_BACKWARD_MSG.format("Solve", "_solve.py:70-131"))

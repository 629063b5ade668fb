"""Solve Function (reference: functions/_solve.py:10-68)."""
from __future__ import annotations

import torch
from torch.autograd import Function

from .. import settings


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
        ctx.representation_tree = representation_tree
        ctx.has_left = has_left
        if has_left:
            left_tensor, right_tensor, *matrix_args = args
        else:
            left_tensor = None
            right_tensor, *matrix_args = args
        orig_right_tensor = right_tensor
        linear_op = representation_tree(*matrix_args)
        ctx.is_vector = right_tensor.ndimension() == 1
        if ctx.is_vector:
            right_tensor = right_tensor.unsqueeze(-1)
        if has_left:  # :48-52
            rhs = torch.cat([left_tensor.mT, right_tensor], -1)
            solves = _solve(linear_op, rhs)
            res = left_tensor @ solves[..., left_tensor.size(-2):]
        else:
            solves = _solve(linear_op, right_tensor)
            res = solves
        if ctx.is_vector:
            res = res.squeeze(-1)
        if has_left:
            ctx.save_for_backward(solves, left_tensor, orig_right_tensor, *matrix_args)
        else:
            ctx.save_for_backward(solves, orig_right_tensor, *matrix_args)
        return res

    @staticmethod
    def backward(ctx, grad_output):
        """Reference :70-131: d/dK (l^T K^-1 r) = -(K^-1 l)(K^-1 r)^T, symmetrised by stacking both orders."""
        if ctx.has_left:
            solves, left_tensor, right_tensor, *matrix_args = ctx.saved_tensors
            left_solves = solves[..., : left_tensor.size(-2)]
            right_solves = solves[..., left_tensor.size(-2):]
        else:
            right_solves, right_tensor, *matrix_args = ctx.saved_tensors
        linear_op = ctx.representation_tree(*matrix_args)
        arg_grads = [None] * len(matrix_args)
        left_grad = right_grad = None
        if not any(ctx.needs_input_grad):
            return tuple([None] * (len(ctx.needs_input_grad)))
        if ctx.is_vector:
            grad_output = grad_output.unsqueeze(-1)
        first_arg = 4 if ctx.has_left else 3
        if not ctx.has_left:
            left_solves = Solve.apply(ctx.representation_tree, False, grad_output, *matrix_args)  # K^-1 grad (:96)
        else:
            left_solves = left_solves @ grad_output  # :113
            if ctx.needs_input_grad[2]:
                left_grad = grad_output @ right_solves.mT  # :116
        if any(ctx.needs_input_grad[first_arg:]):
            arg_grads = linear_op._bilinear_derivative(
                torch.cat([left_solves, right_solves], -1),
                torch.cat([right_solves, left_solves], -1).mul(-0.5),
            )
        if ctx.needs_input_grad[first_arg - 1]:
            right_grad = left_solves.squeeze(-1) if ctx.is_vector else left_solves
        if ctx.has_left:
            return tuple([None, None, left_grad, right_grad] + list(arg_grads))
        return tuple([None, None, right_grad] + list(arg_grads))

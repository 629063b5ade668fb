"""InvQuadLogdet Function: probes -> stacked right-hand side -> preconditioned mBCG with tridiagonal recovery ->
stochastic Lanczos quadrature + inverse quadratic form (reference: functions/_inv_quad_logdet.py:27-161)."""
from __future__ import annotations

import torch
from torch.autograd import Function

from .. import _kernels, settings
from ..utils.stochastic_lq import StochasticLQ


def _draw_probes(precond_lt, num_probes):
    """Probe vectors z ~ N(0, precond_lt), shaped (*batch, N, S), column-normalised, with their norms
    (reference :107-110).  The torch.randn calls replicate the reference's RNG stream: root samples (*b, k, S) first,
    then diagonal samples (S, *b, N) (psd_sum_linear_operator.py:18, _linear_operator.py:2784-2791,
    diag_linear_operator.py:273-277, identity_linear_operator.py:262-266); scaling, summation, transposition and
    normalisation are one fused kernel (lob_probe_assemble)."""
    from ..operators import (DiagLinearOperator, IdentityLinearOperator, PsdSumLinearOperator, RootLinearOperator)

    dtype, device = precond_lt.dtype, precond_lt.device
    if isinstance(precond_lt, IdentityLinearOperator):
        eps = torch.randn(num_probes, *precond_lt.batch_shape, precond_lt.size(-1), dtype=dtype, device=device)
        return _kernels.probe_assemble(None, eps, None)
    if (
        isinstance(precond_lt, PsdSumLinearOperator)
        and len(precond_lt.linear_ops) == 2
        and isinstance(precond_lt.linear_ops[0], RootLinearOperator)
        and isinstance(precond_lt.linear_ops[1], DiagLinearOperator)
    ):
        root_op, diag_op = precond_lt.linear_ops
        L = root_op._root_tensor()
        eps_root = torch.randn(*root_op.batch_shape, L.size(-1), num_probes, dtype=dtype, device=device)
        z_root = _kernels.matmul_nn(L, eps_root)
        eps_diag = torch.randn(num_probes, *diag_op._diag.shape, dtype=dtype, device=device)
        return _kernels.probe_assemble(z_root, eps_diag, diag_op._diag)
    samples = precond_lt.zero_mean_mvn_samples(num_probes)  # (S, *batch, N)
    return _kernels.probe_assemble(None, samples, None)


class InvQuadLogdet(Function):
    @staticmethod
    def forward(ctx, representation_tree, precond_representation_tree, preconditioner, num_precond_args, inv_quad,
                probe_vectors, probe_vector_norms, *args):
        inv_quad_rhs = None
        if inv_quad:
            inv_quad_rhs = args[0]
            args = args[1:]
        if num_precond_args:
            matrix_args = args[:-num_precond_args]
            precond_args = args[-num_precond_args:]
        else:
            matrix_args, precond_args = args, ()

        ctx.representation_tree = representation_tree
        ctx.precond_representation_tree = precond_representation_tree
        ctx.preconditioner = preconditioner
        ctx.inv_quad = inv_quad
        ctx.num_precond_args = num_precond_args
        linear_op = representation_tree(*matrix_args)
        precond_lt = precond_representation_tree(*precond_args)
        dtype, device = linear_op.dtype, linear_op.device
        batch_shape = linear_op.batch_shape
        n = linear_op.matrix_shape[-1]

        if probe_vectors is None or probe_vector_norms is None:  # :78-110
            if settings.deterministic_probes.on():
                raise NotImplementedError(
                    "settings.deterministic_probes is deprecated in the reference and not supported on this path."
                )
            probe_vectors, probe_vector_norms = _draw_probes(precond_lt, settings.num_trace_samples.value())

        num_random_probes = probe_vectors.size(-1)
        rhs_list = [probe_vectors]  # probes FIRST (:118)
        num_inv_quad_solves = 0
        ctx.is_vector = False
        if inv_quad:
            if inv_quad_rhs.ndimension() == 1:
                inv_quad_rhs = inv_quad_rhs.unsqueeze(-1)
                ctx.is_vector = True
            rhs_list.append(inv_quad_rhs)
            num_inv_quad_solves = inv_quad_rhs.size(-1)
        rhs = torch.cat(rhs_list, -1)  # :132
        solves, t_mat = linear_op._solve(rhs, preconditioner, num_tridiag=num_random_probes)  # :133

        logdet_term = torch.zeros(batch_shape, dtype=dtype, device=device)
        inv_quad_term = torch.zeros(batch_shape, dtype=dtype, device=device)
        if settings.skip_logdet_forward.off():  # :140-148
            logdet_term = StochasticLQ.logdet_from_tridiag(t_mat, n)
            if torch.isnan(logdet_term).any().item():  # any NaN in t_mat -> scalar NaN (:141-142)
                logdet_term = torch.tensor(float("nan"), dtype=dtype, device=device)
        if inv_quad:  # :151-153
            inv_quad_term = _kernels.col_dots(solves, num_random_probes, inv_quad_rhs, 0, num_inv_quad_solves)
        ctx.probe_vectors, ctx.probe_vector_norms = probe_vectors, probe_vector_norms
        ctx.num_random_probes, ctx.num_inv_quad_solves = num_random_probes, num_inv_quad_solves
        ctx.save_for_backward(*precond_args, *matrix_args, solves)  # :155-158
        return inv_quad_term, logdet_term

    @staticmethod
    def backward(ctx, inv_quad_grad_output, logdet_grad_output):
        """Reference :163-226.  The probe solves stand in for K^-1 in the log-determinant's gradient; every operator
        gradient is a ``_bilinear_derivative(left, right)`` call (a rank-C outer-product kernel for dense operators),
        the preconditioner's tensors get theirs through ``precond_lt._bilinear_derivative`` (:209-211)."""
        if ctx.num_precond_args:
            precond_args = ctx.saved_tensors[: ctx.num_precond_args]
            matrix_args = ctx.saved_tensors[ctx.num_precond_args: -1]
        else:
            precond_args = []
            matrix_args = ctx.saved_tensors[:-1]
        solves = ctx.saved_tensors[-1]
        linear_op = ctx.representation_tree(*matrix_args)
        precond_lt = ctx.precond_representation_tree(*precond_args)

        if ctx.inv_quad:
            inv_quad_grad_output = inv_quad_grad_output.unsqueeze(-2)  # (*b, 1, R)
        logdet_grad_output = logdet_grad_output.unsqueeze(-1).unsqueeze(-1)  # (*b, 1, 1)

        s = ctx.num_random_probes
        coef = 1.0 / ctx.probe_vectors.size(-1)  # :181
        probe_vector_solves = solves.narrow(-1, 0, s) * (ctx.probe_vector_norms * logdet_grad_output * coef)  # :182-183
        unnormed = ctx.probe_vectors * ctx.probe_vector_norms
        if ctx.preconditioner is not None:  # :187-190: probes now ~ N(0, P^-1)
            precond_probe_vectors = ctx.preconditioner(unnormed)
        else:
            precond_probe_vectors = unnormed

        left_factors_list = [probe_vector_solves]
        right_factors_list = [precond_probe_vectors]
        neg_inv_quad_solves_times_grad_out = None
        if ctx.inv_quad:  # :198-202
            inv_quad_solves = solves.narrow(-1, s, ctx.num_inv_quad_solves)
            neg_inv_quad_solves_times_grad_out = inv_quad_solves * inv_quad_grad_output.neg()
            left_factors_list.append(neg_inv_quad_solves_times_grad_out)
            right_factors_list.append(inv_quad_solves)
        left_factors = torch.cat(left_factors_list, -1)
        right_factors = torch.cat(right_factors_list, -1)
        matrix_arg_grads = linear_op._bilinear_derivative(left_factors, right_factors)  # :206

        precond_arg_grads = precond_lt._bilinear_derivative(  # :209-211
            precond_probe_vectors * (-coef), precond_probe_vectors * logdet_grad_output
        )

        if ctx.inv_quad:  # :213-218
            inv_quad_rhs_grad = neg_inv_quad_solves_times_grad_out * -2.0
            if ctx.is_vector:
                inv_quad_rhs_grad = inv_quad_rhs_grad.squeeze(-1)
            res = [inv_quad_rhs_grad] + list(matrix_arg_grads) + list(precond_arg_grads)
        else:
            res = list(matrix_arg_grads) + list(precond_arg_grads)
        return tuple([None] * 7 + res)

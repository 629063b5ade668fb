"""InvQuadLogdet Function: probes -> stacked right-hand side -> preconditioned mBCG with tridiagonal recovery ->
stochastic Lanczos quadrature + inverse quadratic form (reference: functions/_inv_quad_logdet.py:27-161)."""
from __future__ import annotations

import torch
from torch.autograd import Function

from .. import _kernels, settings
from ..utils.stochastic_lq import StochasticLQ
from ._pivoted_cholesky import _BACKWARD_MSG


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
        if inv_quad:
            if inv_quad_rhs.ndimension() == 1:
                inv_quad_rhs = inv_quad_rhs.unsqueeze(-1)
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
        ctx.mark_non_differentiable(inv_quad_term, logdet_term)
        return inv_quad_term, logdet_term

    @staticmethod
    def backward(ctx, inv_quad_grad_output, logdet_grad_output):
        raise NotImplementedError(_BACKWARD_MSG.format("InvQuadLogdet", "_inv_quad_logdet.py:163-226"))

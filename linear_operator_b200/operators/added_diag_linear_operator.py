"""AddedDiagLinearOperator: ``A + D`` with the pivoted-Cholesky preconditioner
(reference: operators/added_diag_linear_operator.py)."""
from __future__ import annotations

import warnings
from typing import Callable, Optional, Tuple

import torch

from .. import _kernels, settings
from ..utils.warnings import NumericalWarning
from ._linear_operator import LinearOperator
from .dense_linear_operator import DenseLinearOperator
from .diag_linear_operator import ConstantDiagLinearOperator, DiagLinearOperator
from .root_linear_operator import RootLinearOperator
from .sum_linear_operator import PsdSumLinearOperator, SumLinearOperator


class _PreconditionerLogdet(torch.autograd.Function):
    """log det(L L^T + D) as a differentiable function of the pivoted-Cholesky factor and the diagonal.  The reference
    gets this gradient from autograd through its QR (added_diag_linear_operator.py:164-183); here the value comes out
    of the Gram/Cholesky kernels and the gradient is the closed form d/dL = 2 P^-1 L, d/dD = diag(P^-1) with
    P^-1 = D^-1 - Q Q^T applied by the same preconditioner kernels.  For a constant diagonal the reference differentiates
    w.r.t. the ONE value it keeps (``noise.narrow(-2, 0, 1)``, :161), i.e. d/d sigma^2 = tr(P^-1): reproduced, so the
    gradient lands where the reference's does."""

    @staticmethod
    def forward(ctx, precond, L, noise):
        ctx.precond = precond
        ctx.noise_shape = noise.shape
        ctx.save_for_backward(L)
        return precond.logdet.clone()

    @staticmethod
    def backward(ctx, grad):
        (L,) = ctx.saved_tensors
        pre = ctx.precond
        g = grad.reshape(-1)  # (B,)
        B, N = pre.Q.shape[0], pre.N
        grad_L = grad_noise = None
        if ctx.needs_input_grad[1]:
            pinv_l = pre(L.detach())  # P^-1 L through the preconditioner kernels, (*b, N, k)
            grad_L = pinv_l * (2.0 * grad).reshape(*grad.shape, 1, 1)
        if ctx.needs_input_grad[2]:
            rowsq = _kernels.bilinear_diag(pre.Q, pre.Q)  # (B, N): diag(Q Q^T)
            if pre.constant:
                sig = pre.noise[:, 0]
                grad_noise = (g * (N - rowsq.sum(-1)) / sig).reshape(*grad.shape, 1)
            else:
                grad_noise = (g.unsqueeze(-1) * (pre.noise.reciprocal() - rowsq)).reshape(*grad.shape, N)
            if grad_noise.shape != ctx.noise_shape:
                grad_noise = grad_noise.sum_to_size(ctx.noise_shape)
        return None, grad_L, grad_noise


class AddedDiagLinearOperator(SumLinearOperator):
    """``linear_op + diag``; exactly one of the two operands must be a DiagLinearOperator (reference :36-70)."""

    def __init__(self, *linear_ops, preconditioner_override: Optional[Callable] = None):
        linear_ops = list(linear_ops)
        super().__init__(*linear_ops, preconditioner_override=preconditioner_override)
        if len(linear_ops) > 2:
            raise RuntimeError("An AddedDiagLinearOperator can only have two components")
        a, b = self.linear_ops
        if isinstance(a, DiagLinearOperator) and isinstance(b, DiagLinearOperator):
            raise RuntimeError("Trying to lazily add two DiagLinearOperators. Create a single DiagLinearOperator instead.")
        elif isinstance(a, DiagLinearOperator):
            self._diag_tensor, self._linear_op = a, b
        elif isinstance(b, DiagLinearOperator):
            self._diag_tensor, self._linear_op = b, a
        else:
            raise RuntimeError("One of the LinearOperators input to AddedDiagLinearOperator must be a DiagLinearOperator!")
        self.preconditioner_override = preconditioner_override
        self._constant_diag = None
        self._noise = None
        self._piv_chol_self = None
        self._precond_lt = None
        self._precond_logdet_cache = None
        self._q_cache = None  # an _kernels.AddedDiagPreconditioner once built (holds Q and logdet M)

    # ------------------------------------------------------------------ matmul: A X + d (.) X in one pass
    def _matmul(self, rhs):  # :72-76
        if isinstance(self._linear_op, DenseLinearOperator):
            return _kernels.dense_matmul(self._linear_op.tensor, rhs, d=self._diag_tensor._diag)
        fused = getattr(self._linear_op, "_matmul_add_diag", None)
        if fused is not None:
            return fused(rhs, self._diag_tensor._diag)
        out = self._linear_op._matmul(rhs)
        return out.add_(self._diag_tensor._matmul(rhs))

    def _matmul_closure(self):
        if isinstance(self._linear_op, DenseLinearOperator):
            tsr, d = self._linear_op.tensor, self._diag_tensor._diag

            def closure(v):
                return _kernels.dense_matmul(tsr, v, d=d)

            closure.fused = lambda v: _kernels.dense_matmul(tsr, v, d=d, want_dots=True)
            closure.graph_spec = (tsr, d)  # small solves replay as one CUDA graph (settings.cuda_graphs)
            return closure
        fused = getattr(self._linear_op, "_matmul_add_diag", None)
        if fused is not None:  # Kronecker / Toeplitz: <p, A p> comes out of the product's last pass
            d = self._diag_tensor._diag

            def closure(v):
                return fused(v, d)

            closure.fused = lambda v: fused(v, d, want_dots=True)
            return closure
        return self._matmul

    def add_diagonal(self, diag):  # :78-82
        return self.__class__(self._linear_op, self._diag_tensor.add_diagonal(diag))

    def __add__(self, other):  # :84-93
        if isinstance(other, DiagLinearOperator):
            return self.__class__(self._linear_op, self._diag_tensor + other)
        return self.__class__(self._linear_op + other, self._diag_tensor)

    # ------------------------------------------------------------------ preconditioner
    def _preconditioner(self) -> Tuple[Optional[Callable], Optional[LinearOperator], Optional[torch.Tensor]]:
        """Partial pivoted-Cholesky preconditioner M = L L^T + D (reference :95-142): returns
        (closure v -> M^-1 v, PsdSum(Root(L), D), logdet M); (None, None, None) when disabled or too small."""
        if self.preconditioner_override is not None:
            return self.preconditioner_override(self)
        if settings.max_preconditioner_size.value() == 0 or self.size(-1) < settings.min_preconditioning_size.value():
            return None, None, None
        if self._q_cache is None:
            max_iter = settings.max_preconditioner_size.value()
            self._piv_chol_self = self._linear_op.pivoted_cholesky(rank=max_iter)
            self._init_cache()
            # ONE host read for both failure modes: NaNs in the pivoted-Cholesky factor (:126-131) and a non-positive
            # pivot in the k x k factorisation behind Q (the reference's QR would hand back NaNs for it).  The read is
            # issued after the whole build has been queued, so the device keeps working while the host waits.
            bad = torch.isnan(self._piv_chol_self).any() | self._q_cache.info.ne(0).any()
            if bad.item():
                self._q_cache = self._precond_lt = self._precond_logdet_cache = None
                warnings.warn(
                    "NaNs encountered in preconditioner computation. Attempting to continue without preconditioning.",
                    NumericalWarning,
                )
                return None, None, None
        return self._q_cache, self._precond_lt, self._precond_logdet_cache

    def _init_cache(self):  # :144-184
        diag = self._diag_tensor._diagonal()
        if isinstance(self._diag_tensor, ConstantDiagLinearOperator):
            self._constant_diag = True
        else:
            # the reference decides "constant" at run time by comparing with the first element (:149-150)
            self._constant_diag = bool(torch.equal(diag, diag[..., :1].expand_as(diag)))
        self._noise = diag[..., :1] if self._constant_diag else diag
        with torch.no_grad():
            self._q_cache = _kernels.AddedDiagPreconditioner(self._piv_chol_self, diag, self._constant_diag)
        self._precond_logdet_cache = self._q_cache.logdet
        if torch.is_grad_enabled() and (self._piv_chol_self.requires_grad or diag.requires_grad):
            # logdet_P is added to the estimate (operators/_linear_operator.py:1799-1800) and is part of the graph
            self._precond_logdet_cache = _PreconditionerLogdet.apply(self._q_cache, self._piv_chol_self, self._noise)
        self._precond_lt = PsdSumLinearOperator(RootLinearOperator(self._piv_chol_self), self._diag_tensor)  # :159

    def _diagonal(self):
        return self._linear_op._diagonal() + self._diag_tensor._diagonal()


__all__ = ["AddedDiagLinearOperator"]

"""Shifted MINRES -- host side (reference: utils/minres.py:10-282): solves (K * value + shift_q I) x_q = b for all
shifts at once; same signature, shapes and stop rule as the reference.  The operator and the preconditioner stay closures;
an iteration is the closure product, two deterministic column reductions and three fused launches (csrc/minres.cu), and
the host reads one number every 10th iteration (the reference's own convergence test, :177-182)."""
from __future__ import annotations

import torch

from .. import _kernels, _lib, settings
from .._lib import check, dt, ptr, require_cuda, stream


@_lib.device_guard
def minres(matmul_closure, rhs, eps=1e-25, shifts=None, value=None, max_iter=None, preconditioner=None):
    """Returns the solves, shaped like ``rhs`` with a leading shift dimension (squeezed for a single shift)."""
    if torch.is_tensor(matmul_closure):
        mat = matmul_closure

        def matmul_closure(v):
            return _kernels.dense_matmul(mat, v)

    mm_ = matmul_closure
    require_cuda(rhs, shifts)
    lib = _lib.load()
    if shifts is None:
        shifts = torch.tensor(0.0, dtype=rhs.dtype, device=rhs.device)
    squeeze = False
    if rhs.dim() == 1:
        rhs = rhs.unsqueeze(-1)
        squeeze = True

    # one-time set-up on small / single passes (:47-50,:60-77)
    rhs_norm = rhs.norm(2, dim=-2, keepdim=True)
    rhs_is_zero = rhs_norm.lt(1e-10)
    rhs_norm = rhs_norm.masked_fill_(rhs_is_zero, 1)
    rhs = rhs.div(rhs_norm)
    if max_iter is None:
        max_iter = settings.max_cg_iterations.value()
    max_iter = min(max_iter, rhs.size(-2) + 1)

    prod = mm_(rhs)  # fixes the broadcast batch shape (:58)
    full_shape = prod.shape
    batch_shape = full_shape[:-2]
    N, C = full_shape[-2:]
    B = 1
    for sdim in batch_shape:
        B *= int(sdim)
    n_shift_dims = shifts.dim()
    shifts_p = shifts.reshape(*shifts.shape, *([1] * (prod.dim() - n_shift_dims + 1)))  # _pad_with_singletons (:62)
    Q = shifts_p.shape[0]
    shifts_flat = shifts_p.expand(Q, *batch_shape, 1, 1).reshape(Q, B).to(rhs.dtype).contiguous()

    dev, dty = rhs.device, rhs.dtype
    st = stream(rhs)
    zeros = lambda *shape: torch.zeros(*shape, dtype=dty, device=dev)  # noqa: E731
    solution = zeros(Q, B, N, C)
    z2 = zeros(B, N, C)
    z1 = rhs.expand(full_shape).reshape(B, N, C).clone()
    if preconditioner is None:
        q1 = z1.clone()
    else:
        q1 = preconditioner(z1.reshape(full_shape)).reshape(B, N, C).contiguous()
    beta_prev = _kernels.col_dots(z1, 0, q1, 0, C).sqrt_()  # (B, C)  (:69)
    z1.div_(beta_prev.unsqueeze(-2))
    q1.div_(beta_prev.unsqueeze(-2))
    beta_curr = torch.empty_like(beta_prev)
    cos2, sin2 = torch.ones(Q, B, C, dtype=dty, device=dev), zeros(Q, B, C)
    cos1, sin1 = torch.ones_like(cos2), torch.zeros_like(sin2)
    cos_c, sin_c = torch.empty_like(cos2), torch.empty_like(cos2)
    sub, subsub, diag = torch.empty_like(cos2), torch.empty_like(cos2), torch.empty_like(cos2)
    search2, search1, search_c = zeros(Q, B, N, C), zeros(Q, B, N, C), torch.empty(Q, B, N, C, dtype=dty, device=dev)
    scale_prev = beta_prev.unsqueeze(0).repeat(Q, 1, 1)
    scale_curr = torch.empty_like(scale_prev)

    if settings.verbose_linalg.on():
        settings.verbose_linalg.logger.debug(
            f"Running MINRES on a {rhs.shape} RHS for {max_iter} iterations (tol={settings.minres_tolerance.value()}). "
            f"Output: {(Q, *full_shape)}."
        )

    d = dt(rhs)
    for i in range(max_iter + 2):
        prod = mm_(q1.reshape(full_shape))
        if value is not None:
            prod = prod * value if not prod.is_contiguous() else prod.mul_(value)
        prod = prod.reshape(B, N, C).contiguous()
        alpha = _kernels.col_dots(prod, 0, q1, 0, C)  # (:132-133)
        check(lib.lob_minres_z(d, B, N, C, ptr(prod), ptr(z1), ptr(z2), ptr(alpha), ptr(beta_prev), st), "lob_minres_z")
        z = prod
        if preconditioner is None:
            q = z  # the reference clones; the update kernel normalises the shared buffer once
            bsq = _kernels.col_dots(z, 0, z, 0, C)
        else:
            q = preconditioner(z.reshape(full_shape)).reshape(B, N, C).contiguous()
            bsq = _kernels.col_dots(z, 0, q, 0, C)
        check(
            lib.lob_minres_scalars(d, Q, B, C, ptr(shifts_flat), ptr(alpha), ptr(beta_prev), ptr(bsq), ptr(beta_curr),
                                   ptr(cos2), ptr(sin2), ptr(cos1), ptr(sin1), ptr(cos_c), ptr(sin_c), ptr(scale_prev),
                                   ptr(scale_curr), ptr(sub), ptr(subsub), ptr(diag), float(eps), st),
            "lob_minres_scalars",
        )
        check(
            lib.lob_minres_update(d, Q, B, N, C, ptr(z), ptr(q), ptr(beta_curr), ptr(q1), ptr(search1), ptr(search2),
                                  ptr(search_c), ptr(solution), ptr(sub), ptr(subsub), ptr(diag), ptr(scale_prev), st),
            "lob_minres_update",
        )
        if (i + 1) % 10 == 0:  # :177-182
            upd = _kernels.col_dots(search_c.reshape(Q * B, N, C), 0, search_c.reshape(Q * B, N, C), 0, C).sqrt_()
            upd.mul_(scale_prev.reshape(Q * B, C).abs())
            sol = _kernels.col_dots(solution.reshape(Q * B, N, C), 0, solution.reshape(Q * B, N, C), 0, C).sqrt_()
            conv = upd.div_(sol).mean().item()
            if conv < settings.minres_tolerance.value():
                break
        # rotate (:185-195)
        z2, z1 = z1, z
        q1 = q
        beta_prev, beta_curr = beta_curr, beta_prev
        cos2, cos1, cos_c = cos1, cos_c, cos2
        sin2, sin1, sin_c = sin1, sin_c, sin2
        search2, search1, search_c = search1, search_c, search2
        scale_prev, scale_curr = scale_curr, scale_prev

    solution = solution.reshape(Q, *full_shape)
    solution.masked_fill_(rhs_is_zero, 0)
    if squeeze:
        solution = solution.squeeze(-1)
        rhs_norm = rhs_norm.squeeze(-1)
    if shifts.numel() == 1:
        solution = solution.squeeze(0)
    return solution.mul_(rhs_norm)

"""Modified batched conjugate gradients -- host side.

Mirror of the reference's ``linear_operator.utils.linear_cg`` (utils/linear_cg.py:98-359): same signature, argument
meaning, return values, RuntimeError / NumericalWarning behaviour.  The operator matmul and the preconditioner stay
closures (the reference's plug-in points); everything else the reference does per iteration (~40 ATen launches and two
host synchronisations, :250-332) is three fused launches in ``csrc/cg_kernels.cu`` and no host round trip until the
first iteration at which the reference's own stop rule could fire.
"""
from __future__ import annotations

import warnings

import torch

from .. import _lib, settings
from .._lib import CgParams, CgStatus, check, dt, ptr, require_cuda, stream, workspace
from .warnings import NumericalWarning


def _default_preconditioner(x):
    return x.clone()


@_lib.device_guard
def linear_cg(
    matmul_closure,
    rhs,
    n_tridiag=0,
    tolerance=None,
    eps=1e-10,
    stop_updating_after=1e-10,
    max_iter=None,
    max_tridiag_iter=None,
    initial_guess=None,
    preconditioner=None,
    _skip_initial_matmul=False,
):
    """Solves ``lhs result = rhs`` for symmetric positive definite ``lhs`` given as a matmul closure.

    Args and returns as in the reference (utils/linear_cg.py:110-132): returns ``result`` or, when ``n_tridiag > 0``,
    ``(result, t_mat)`` with ``t_mat`` of shape ``(n_tridiag, *batch, T, T)``.

    ``matmul_closure`` may carry an attribute ``fused(p) -> (Ap, dots, n_parts)`` (set by the dense operators) whose
    partial ``<p, Ap>`` sums come out of the matmul epilogue; plain callables and tensors work as in the reference.
    ``_skip_initial_matmul`` (private, used by ``LinearOperator._solve``): with no initial guess the reference still
    multiplies the operator with a zero vector (:186); operators of this package promise ``A 0 = 0`` with the right
    batch shape, so that product is skipped and the NaN check (:199) runs on the first real product instead.
    """
    is_vector = rhs.ndimension() == 1
    if is_vector:
        rhs = rhs.unsqueeze(-1)

    # defaults (:139-156)
    if max_iter is None:
        max_iter = settings.max_cg_iterations.value()
    if max_tridiag_iter is None:
        max_tridiag_iter = settings.max_lanczos_quadrature_iterations.value()
    have_guess = initial_guess is not None
    if have_guess and initial_guess.ndimension() == 1:
        initial_guess = initial_guess.unsqueeze(-1)
    if tolerance is None:
        tolerance = settings.cg_tolerance.value()
    precond = preconditioner is not None

    if max_tridiag_iter > max_iter:  # :159-160
        raise RuntimeError("Getting a tridiagonalization larger than the number of CG iterations run is not possible!")

    fused = getattr(matmul_closure, "fused", None)
    if torch.is_tensor(matmul_closure):  # :163-166
        from .. import _kernels

        mat = matmul_closure

        def matmul_closure(v):
            return _kernels.dense_matmul(mat, v)

        def fused(v):
            return _kernels.dense_matmul(mat, v, want_dots=True)

    elif not callable(matmul_closure):
        raise RuntimeError("matmul_closure must be a tensor, or a callable object!")

    require_cuda(rhs, initial_guess)
    lib = _lib.load()

    num_rows = rhs.size(-2)
    n_iter = min(max_iter, num_rows) if settings.terminate_cg_by_size.on() else max_iter  # :170
    n_tridiag_iter = min(max_tridiag_iter, num_rows)  # :171

    if settings.verbose_linalg.on():
        settings.verbose_linalg.logger.debug(
            f"Running CG on a {rhs.shape} RHS for {n_iter} iterations (tol={tolerance}). Output: {rhs.shape}."
        )

    # The operator's batch shape may be larger than the rhs's (:186-190).  Without the initial product we rely on the
    # caller (LinearOperator._solve) to have expanded rhs; foreign closures take the reference's exact route.
    ax0 = None
    if have_guess or not _skip_initial_matmul:
        guess = initial_guess if have_guess else torch.zeros_like(rhs)
        # reference: rhs_norm scaling first, then A x0 -- the scaling commutes with A column-wise, so multiply the
        # unscaled guess once the norms are known; we need the norms on device first, so do it after setup below.
        probe_shape = torch.broadcast_shapes(rhs.shape, guess.shape)
        rhs = rhs.expand(probe_shape)
        initial_guess = guess.expand(probe_shape)
        have_guess = True

    rhs_c = rhs.contiguous()
    batch_shape = rhs_c.shape[:-2]
    N, C = rhs_c.shape[-2:]
    B = 1
    for s in batch_shape:
        B *= s

    def make_params(B_):
        return CgParams(B_, N, C, dt(rhs_c), int(n_tridiag), int(n_tridiag_iter), int(max_iter), int(n_iter),
                        1 if precond else 0, float(tolerance), float(eps), float(stop_updating_after))

    p = make_params(B)
    dev = rhs_c.device
    st = stream(rhs_c)
    ws = workspace(lib.lob_cg_workspace_bytes(ctypes_byref(p)), dev)
    rhs_n = torch.empty_like(rhs_c)
    x = torch.empty_like(rhs_c)
    t_mat = None
    if n_tridiag:
        t_mat = torch.empty(n_tridiag, *batch_shape, n_tridiag_iter, n_tridiag_iter, dtype=rhs_c.dtype, device=dev)
    x0_c = initial_guess.contiguous() if have_guess else None
    check(lib.lob_cg_setup(ctypes_byref(p), ptr(ws), ptr(rhs_c), ptr(x0_c), ptr(rhs_n), ptr(x), ptr(t_mat), st),
          "lob_cg_setup")

    if have_guess:
        ax0 = matmul_closure(x)  # x = x0 / ||rhs||  (:183,186)
        if ax0.shape != x.shape:
            # the operator is batched more widely than the rhs: restart with everything expanded (:187-190)
            full = ax0.shape
            res = linear_cg(
                matmul_closure, rhs.expand(*full[:-2], N, C),
                n_tridiag=n_tridiag, tolerance=tolerance, eps=eps, stop_updating_after=stop_updating_after,
                max_iter=max_iter, max_tridiag_iter=max_tridiag_iter,
                initial_guess=initial_guess.expand(*full[:-2], N, C), preconditioner=preconditioner,
            )
            if not is_vector:
                return res
            return (res[0].squeeze(-1), res[1]) if n_tridiag else res.squeeze(-1)
        ax0 = ax0.contiguous()

    r = torch.empty_like(rhs_c)
    check(lib.lob_cg_residual_init(ctypes_byref(p), ptr(ws), ptr(rhs_n), ptr(ax0), ptr(r), st), "lob_cg_residual_init")
    del ax0, rhs_n

    status = CgStatus()

    def poll():
        check(lib.lob_cg_poll_sync(ctypes_byref(p), ptr(ws), ctypes_byref(status), st), "lob_cg_poll_sync")
        if status.nan_detected:  # :199-200
            raise RuntimeError("NaNs encountered when trying to perform matrix-vector multiplication")
        return status.stop

    if have_guess:
        poll()  # the reference checks for NaNs right here (:199); with a user guess we keep that timing

    precond_fused = getattr(preconditioner, "fused", None) if precond else None

    def apply_precond(res):
        """z = M^-1 r (:213,:268); preconditioners with a fused epilogue also hand back the <r, z> partial sums"""
        if precond_fused is not None:
            zz, rz_parts, n_rz = precond_fused(res)
            return zz.contiguous(), rz_parts, n_rz
        return preconditioner(res).contiguous(), None, 0

    if precond:
        z, rz_parts, n_rz = apply_precond(r)
    else:
        z, rz_parts, n_rz = r, None, 0
    pvec = torch.empty_like(r)
    check(lib.lob_cg_direction_init(ctypes_byref(p), ptr(ws), ptr(r), ptr(z), ptr(pvec), ptr(rz_parts), n_rz, st),
          "lob_cg_direction_init")

    # first iteration index at which the reference's stop rule (:302-306) can fire
    first_stop = min(10, max_iter - 1)
    if n_tridiag:
        first_stop = max(first_stop, min(n_tridiag_iter, max_iter - 1))

    polled = False
    for k in range(n_iter):
        if fused is not None:
            ap, dots, n_parts = fused(pvec)
        else:
            ap, dots, n_parts = matmul_closure(pvec), None, 0
        if ap.shape != pvec.shape:
            raise RuntimeError(
                f"matmul_closure returned shape {tuple(ap.shape)} for an input of shape {tuple(pvec.shape)}; expand "
                "the right-hand side to the operator's batch shape (LinearOperator._solve does this)."
            )
        ap = ap.contiguous()
        check(lib.lob_cg_step_xr(ctypes_byref(p), ptr(ws), k, ptr(ap), ptr(pvec), ptr(x), ptr(r), ptr(dots), n_parts,
                                 st), "lob_cg_step_xr")
        if precond:
            z, rz_parts, n_rz = apply_precond(r)
        else:
            z, rz_parts, n_rz = None, None, 0
        check(lib.lob_cg_step_p(ctypes_byref(p), ptr(ws), k, ptr(z), ptr(r), ptr(pvec), ptr(t_mat), ptr(rz_parts), n_rz,
                                st), "lob_cg_step_p")
        polled = False
        if k == 0 or k >= first_stop or k == n_iter - 1:
            polled = True
            if poll():
                break
    if not polled:
        poll()

    check(lib.lob_cg_finish(ctypes_byref(p), ptr(ws), ptr(x), st), "lob_cg_finish")  # :335

    if not status.tolerance_reached and n_iter > 0 and status.iterations > 0:  # :337-347
        warnings.warn(
            "CG terminated in {} iterations with average residual norm {}"
            " which is larger than the tolerance of {} specified by"
            " linear_operator.settings.cg_tolerance."
            " If performance is affected, consider raising the maximum number of CG iterations by running code in"
            " a linear_operator.settings.max_cg_iterations(value) context.".format(
                status.iterations, status.residual_norm_mean, tolerance
            ),
            NumericalWarning,
        )

    result = x
    if is_vector:
        result = result.squeeze(-1)
    if n_tridiag:  # :352-357
        last = status.last_tridiag_iter + 1
        return result, t_mat[..., :last, :last].contiguous()
    return result


def ctypes_byref(obj):
    import ctypes

    return ctypes.byref(obj)

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


# ------------------------------------------------------------------------------------------------------------
# CUDA-graph replay of small dense solves (settings.cuda_graphs)
# ------------------------------------------------------------------------------------------------------------
_GRAPHS = {}


class _GraphedDenseCg:
    """The launches of one un-preconditioned mBCG solve on a dense (+ diagonal) operator -- setup, initial residual and
    direction, and the first ``n_graph`` iterations (every iteration before the reference's stop rule can fire) --
    captured once into a CUDA graph over static buffers.  The kernels are the eager path's, in the same order, so the
    numbers are bit-identical; what disappears is ~4 host launches per iteration."""

    def __init__(self, A, d, rhs_c, p, n_graph, n_tridiag):
        from .. import _kernels

        lib = _lib.load()
        dev = rhs_c.device
        self.p, self.n_graph = p, n_graph
        self.A = torch.empty(A.shape, dtype=A.dtype, device=dev)
        self.d = None
        self.d_const = False
        if d is not None:
            self.d_const = d.shape[-1] == 1 or d.stride(-1) == 0
            self.d = torch.empty(*d.shape[:-1], 1 if self.d_const else d.shape[-1], dtype=d.dtype, device=dev)
        self.rhs = torch.empty_like(rhs_c)
        self.ws = workspace(lib.lob_cg_workspace_bytes(ctypes_byref(p)), dev)
        self.rhs_n = torch.empty_like(rhs_c)
        self.x = torch.empty_like(rhs_c)
        self.r = torch.empty_like(rhs_c)
        self.pvec = torch.empty_like(rhs_c)
        self.t_mat = None
        if n_tridiag:
            self.t_mat = torch.empty(n_tridiag, *rhs_c.shape[:-2], p.n_tridiag_iter, p.n_tridiag_iter, dtype=rhs_c.dtype,
                                     device=dev)
        self._kernels = _kernels
        self.load(A, d, rhs_c)
        # warm-up on a side stream (lazy module / attribute initialisation must not happen inside the capture)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            self.body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.body()

    def load(self, A, d, rhs_c):
        self.A.copy_(A)
        if self.d is not None:
            self.d.copy_(d[..., :1] if self.d_const else d)
        self.rhs.copy_(rhs_c)

    def iterate(self, k):
        lib, p, st = _lib.load(), self.p, stream(self.rhs)
        ap, dots, n_parts = self._kernels.dense_matmul(self.A, self.pvec, d=self.d, want_dots=True)
        check(lib.lob_cg_step_xr(ctypes_byref(p), ptr(self.ws), k, ptr(ap), ptr(self.pvec), ptr(self.x), ptr(self.r),
                                 ptr(dots), n_parts, st), "lob_cg_step_xr")
        check(lib.lob_cg_step_p(ctypes_byref(p), ptr(self.ws), k, None, ptr(self.r), ptr(self.pvec), ptr(self.t_mat),
                                None, 0, st), "lob_cg_step_p")

    def body(self):
        lib, p, st = _lib.load(), self.p, stream(self.rhs)
        check(lib.lob_cg_setup(ctypes_byref(p), ptr(self.ws), ptr(self.rhs), None, ptr(self.rhs_n), ptr(self.x),
                               ptr(self.t_mat), st), "lob_cg_setup")
        check(lib.lob_cg_residual_init(ctypes_byref(p), ptr(self.ws), ptr(self.rhs_n), None, ptr(self.r), st),
              "lob_cg_residual_init")
        check(lib.lob_cg_direction_init(ctypes_byref(p), ptr(self.ws), ptr(self.r), ptr(self.r), ptr(self.pvec), None, 0,
                                        st), "lob_cg_direction_init")
        for k in range(self.n_graph):
            self.iterate(k)


def _graph_eligible(spec, rhs, precond, have_guess, skip_initial):
    if spec is None or precond or have_guess or not skip_initial or settings.cuda_graphs.off():
        return False
    A, d = spec
    if not (torch.is_tensor(A) and A.is_cuda and rhs.is_cuda) or A.dtype != rhs.dtype:
        return False
    if A.shape[-1] != A.shape[-2] or A.shape[:-2] != rhs.shape[:-2]:
        return False
    if A.numel() * A.element_size() > settings.cuda_graphs.max_operator_bytes:
        return False
    if rhs.numel() * rhs.element_size() > settings.cuda_graphs.max_rhs_bytes:
        return False
    return not torch.cuda.is_current_stream_capturing()


def _warn_not_converged(status, n_iter, tolerance):
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


def _graphed_dense_solve(spec, rhs_c, p, n_iter, n_tridiag, first_stop, tolerance):
    A, d = spec
    lib = _lib.load()
    n_graph = min(n_iter, first_stop + 1)
    key = (rhs_c.device.index, rhs_c.dtype, tuple(A.shape), None if d is None else (tuple(d.shape), d.stride(-1) == 0),
           tuple(rhs_c.shape), n_graph, p.n_tridiag, p.n_tridiag_iter, p.max_iter, p.n_iter, p.tolerance, p.eps,
           p.stop_updating_after)
    g = _GRAPHS.get(key)
    if g is None:
        if len(_GRAPHS) >= 16:  # bounded cache of static buffers
            _GRAPHS.pop(next(iter(_GRAPHS)))
        g = _GRAPHS[key] = _GraphedDenseCg(A, d, rhs_c, p, n_graph, n_tridiag)
    else:
        g.load(A, d, rhs_c)
    g.graph.replay()
    status = CgStatus()
    st = stream(rhs_c)

    def poll():
        check(lib.lob_cg_poll_sync(ctypes_byref(g.p), ptr(g.ws), ctypes_byref(status), st), "lob_cg_poll_sync")
        if status.nan_detected:
            raise RuntimeError("NaNs encountered when trying to perform matrix-vector multiplication")
        return status.stop

    stopped = poll()
    k = n_graph
    while not stopped and k < n_iter:  # the stop rule did not fire inside the captured part: continue launch by launch
        g.iterate(k)
        stopped = poll()
        k += 1
    check(lib.lob_cg_finish(ctypes_byref(g.p), ptr(g.ws), ptr(g.x), st), "lob_cg_finish")
    _warn_not_converged(status, n_iter, tolerance)
    x = g.x.clone()
    if n_tridiag:
        last = status.last_tridiag_iter + 1
        return x, g.t_mat[..., :last, :last].contiguous()
    return x, None


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

    # first iteration index at which the reference's stop rule (:302-306) can fire
    first_stop = min(10, max_iter - 1)
    if n_tridiag:
        first_stop = max(first_stop, min(n_tridiag_iter, max_iter - 1))

    graph_spec = getattr(matmul_closure, "graph_spec", None)
    if n_iter > 0 and _graph_eligible(graph_spec, rhs_c, precond, have_guess, _skip_initial_matmul):
        result, t_mat = _graphed_dense_solve(graph_spec, rhs_c, p, n_iter, n_tridiag, first_stop, tolerance)
        if is_vector:
            result = result.squeeze(-1)
        return (result, t_mat) if n_tridiag else result
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

    _warn_not_converged(status, n_iter, tolerance)

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

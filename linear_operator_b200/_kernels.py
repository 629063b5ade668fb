"""Typed torch-tensor wrappers over the C ABI (``include/lob_b200.h``).  Shape checks live here; arithmetic does not.

Every function takes CUDA tensors, flattens the leading batch dimensions to one ``B`` and launches on the current
stream of the tensors' device.  Nothing here falls back to PyTorch math.
"""
from __future__ import annotations

import ctypes
import math
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import check, device_guard, dt, ptr, require_cuda, stream, workspace


# bench.py sets this to a list to collect (start, end) CUDA events around every dense operator matmul launch, recorded
# on the launching stream (roofline.achieved is derived from them).  None = no profiling.
PROFILE_MATMUL = None


def _flat3(t: torch.Tensor) -> torch.Tensor:
    """(*batch, R, C) -> contiguous (B, R, C)"""
    t = t.contiguous()
    return t.reshape(-1, t.shape[-2], t.shape[-1])


def _numel(shape) -> int:
    return int(math.prod(shape)) if len(shape) else 1


def _diag_args(d: Optional[torch.Tensor], batch_shape, n: int):
    """Normalises a diagonal given as (*batch, N) or (*batch, 1) (constant) to (tensor, batch_stride, elem_stride).
    A stride-0 expanded diagonal (ConstantDiagLinearOperator._diag, diag_linear_operator.py:346-350) is passed through
    without materialising it."""
    if d is None:
        return None, 0, 0
    B = _numel(batch_shape)
    if d.shape[-1] == 1 or (d.dim() >= 1 and d.stride(-1) == 0):
        base = d[..., :1]
        base = base.expand(*batch_shape, 1).reshape(B, 1).contiguous() if base.numel() != B else base.reshape(B, 1).contiguous()
        return base, 1, 0
    full = d.expand(*batch_shape, n).reshape(B, n).contiguous()
    return full, n, 1


# ------------------------------------------------------------------------------------------------------------
# matmuls
# ------------------------------------------------------------------------------------------------------------
def dense_matmul(A: torch.Tensor, X: torch.Tensor, d: Optional[torch.Tensor] = None, want_dots: bool = False,
                 E: Optional[torch.Tensor] = None, alpha: Optional[torch.Tensor] = None):
    """Y = alpha * (A X) + d (.) E with E = X by default; A (*ba, M, K), X (*b, K, C), E (*b, M, C), alpha (B,) or
    None.  Returns Y or (Y, dots, n_parts) where dots are the per-row-tile partial sums of E * Y."""
    require_cuda(A, X, d, E, alpha)
    lib = _lib.load()
    M, K = A.shape[-2:]
    if X.shape[-2] != K:
        raise RuntimeError(f"Size mismatch: operator is {tuple(A.shape)}, right-hand side is {tuple(X.shape)}")
    batch_shape = torch.broadcast_shapes(A.shape[:-2], X.shape[:-2])
    B = _numel(batch_shape)
    C = X.shape[-1]
    Xf = _flat3(X.expand(*batch_shape, K, C))
    if _numel(A.shape[:-2]) == 1:
        Af = A.reshape(1, M, K)
        if Af.stride(-1) != 1 or Af.stride(-2) < K:
            Af = Af.contiguous()
        a_bs = 0
    else:
        Af = _flat3(A.expand(*batch_shape, M, K))
        a_bs = Af.stride(0)
    lda = Af.stride(-2)
    Y = torch.empty(B, M, C, dtype=X.dtype, device=X.device)
    dd, d_bs, d_st = _diag_args(d, batch_shape, M)
    dots = None
    n_parts = int(lib.lob_dense_matmul_parts(M))
    # small fp64 problems (BASELINE config 1) take a row-per-warp kernel without fused <X, Y> partial sums
    # (csrc/matmul_simt.cu, k_matmul_rows: same rule there); linear_cg then adds its own dot pass
    fused_dots = not (X.dtype == torch.float64 and B * M <= 8192 and K <= 8192 and C <= 64 and E is None
                      and alpha is None)
    if want_dots and fused_dots:
        dots = torch.empty(B, n_parts, C, dtype=torch.float64, device=X.device)
    # scratch of the streaming tensor-core kernel (tf32 split of X^T); torch's caching allocator makes this free
    ws_bytes = int(lib.lob_dense_matmul_workspace_bytes(dt(X), B, M, K, C))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=X.device) if ws_bytes else None
    prof = PROFILE_MATMUL
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    if E is None and alpha is None:
        check(
            lib.lob_dense_matmul(dt(X), B, M, K, C, ptr(Af), lda, a_bs, ptr(Xf), ptr(Y), ptr(dd), d_bs, d_st,
                                 ptr(dots), ptr(ws), ws_bytes, stream(X)),
            "lob_dense_matmul",
        )
    else:
        Ef = None if E is None else _flat3(E.expand(*batch_shape, M, C))
        al = None if alpha is None else alpha.reshape(-1).contiguous()
        check(
            lib.lob_dense_matmul_ex(dt(X), B, M, K, C, ptr(Af), lda, a_bs, ptr(Xf), ptr(Y), ptr(Ef), ptr(al),
                                    0 if (al is None or al.numel() == 1) else 1, ptr(dd), d_bs, d_st, ptr(dots),
                                    ptr(ws), ws_bytes, stream(X)),
            "lob_dense_matmul_ex",
        )
    if prof is not None:
        e1.record()
        prof.append((e0, e1, M == K))  # square = the operator matmul itself
    Y = Y.reshape(*batch_shape, M, C)
    if want_dots:
        return Y, dots, (n_parts if dots is not None else 0)
    return Y


def gemm3x(A: torch.Tensor, B: torch.Tensor, trans_a: bool = False, trans_b: bool = False, alpha: float = 1.0,
           row_alpha: Optional[torch.Tensor] = None, E: Optional[torch.Tensor] = None,
           row_beta: Optional[torch.Tensor] = None, splits: int = 0, out_dtype: Optional[torch.dtype] = None,
           a_div: int = 1, b_div: int = 1, batch: Optional[int] = None,
           store_transposed: bool = False) -> Optional[torch.Tensor]:
    """D[b] = ra * op(A[b // a_div]) op(B[b // b_div]) + rb * E[b] on the tensor cores (3xTF32, csrc/gemm3x.cu), BLAS-style:
    A is stored (nb_a, M, K) -- (nb_a, K, M) with trans_a --, B is stored (nb_b, K, N) -- (nb_b, N, K) with trans_b;
    the last dimension of either may be a strided view (leading dimension = stride of dim -2).  row_alpha / row_beta are
    (batch, M) per-row factors.  ``store_transposed``: the result comes back as (batch, N, M) = D^T, E is given in that
    layout too and the factors are (batch, N) -- choose the roles so that M is the memory-contiguous index of the
    result (M = the tensor core's lane index: every store is then a full line).  Returns None when the layout is not TMA-addressable (caller takes the CUDA-core
    kernels); fp32 only."""
    require_cuda(A, B, row_alpha, E, row_beta)
    if A.dtype != torch.float32 or B.dtype != torch.float32:
        return None
    lib = _lib.load()

    def as3(t):
        t = t if t.dim() == 3 else t.reshape(1, *t.shape[-2:]) if t.dim() == 2 else t.reshape(-1, *t.shape[-2:])
        if t.stride(-1) != 1 or t.stride(-2) < t.shape[-1] or (t.shape[0] > 1 and t.stride(0) < t.shape[-2] * t.stride(-2)):
            t = t.contiguous()
        return t

    A3, B3 = as3(A), as3(B)
    M, K = (A3.shape[2], A3.shape[1]) if trans_a else (A3.shape[1], A3.shape[2])
    Kb, N = (B3.shape[2], B3.shape[1]) if trans_b else (B3.shape[1], B3.shape[2])
    if K != Kb:
        raise RuntimeError(f"Size mismatch in gemm3x: {tuple(A.shape)} (trans={trans_a}) x {tuple(B.shape)} (trans={trans_b})")
    nb = batch if batch is not None else max(A3.shape[0] * a_div, B3.shape[0] * b_div)
    out_dtype = out_dtype or torch.float32
    want_splits = int(lib.lob_gemm3x_splits(nb, M, N, K, int(splits)))
    if out_dtype == torch.float64 and want_splits == 1:
        want_splits = int(lib.lob_gemm3x_splits(nb, M, N, K, 2))
        if want_splits == 1:
            return None
        splits = 2
    rows, cols = (N, M) if store_transposed else (M, N)  # stored shape of D and E
    D = torch.empty(nb, rows, cols, dtype=out_dtype, device=A.device)
    ws_bytes = int(lib.lob_gemm3x_workspace_bytes(nb, M, N, K, int(splits)))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=A.device) if ws_bytes else None
    ra = None if row_alpha is None else row_alpha.to(torch.float32).expand(nb, rows).contiguous()
    rb = None if row_beta is None else row_beta.to(torch.float32).expand(nb, rows).contiguous()
    E3 = None
    if E is not None:
        E3 = E.reshape(nb, rows, cols)
        if E3.stride(-1) != 1:
            E3 = E3.contiguous()
    status = lib.lob_gemm3x(
        nb, M, N, K, ptr(A3), 1 if trans_a else 0, A3.stride(1), A3.stride(0) if A3.shape[0] > 1 else 0, int(a_div),
        ptr(B3), 0 if trans_b else 1, B3.stride(1), B3.stride(0) if B3.shape[0] > 1 else 0, int(b_div),
        ptr(D), _lib._DT[out_dtype], cols, M * N, float(alpha), ptr(ra), rows, ptr(E3),
        0 if E3 is None else E3.stride(1), 0 if E3 is None else E3.stride(0), ptr(rb), rows,
        1 if store_transposed else 0, int(splits), ptr(ws), ws_bytes, stream(A))
    if status == _lib.UNSUPPORTED:
        return None
    check(status, "lob_gemm3x")
    return D


def matmul_nn(A: torch.Tensor, X: torch.Tensor) -> torch.Tensor:
    """Y = A X for small-K products (Q t, L eps, U w).  A (*ba, M, K), X (*bx, K, C), batches broadcast."""
    require_cuda(A, X)
    lib = _lib.load()
    M, K = A.shape[-2:]
    C = X.shape[-1]
    if X.shape[-2] != K:
        raise RuntimeError(f"Size mismatch: {tuple(A.shape)} @ {tuple(X.shape)}")
    batch_shape = torch.broadcast_shapes(A.shape[:-2], X.shape[:-2])
    B = _numel(batch_shape)
    if _numel(A.shape[:-2]) == 1:
        Af, a_bs = A.reshape(1, M, K).contiguous(), 0
    else:
        Af = _flat3(A.expand(*batch_shape, M, K))
        a_bs = M * K
    if _numel(X.shape[:-2]) == 1:
        Xf, x_bs = X.reshape(1, K, C).contiguous(), 0
    else:
        Xf = _flat3(X.expand(*batch_shape, K, C))
        x_bs = K * C
    Y = torch.empty(B, M, C, dtype=X.dtype, device=X.device)
    check(lib.lob_matmul_nn(dt(X), B, M, K, C, ptr(Af), K, a_bs, ptr(Xf), x_bs, ptr(Y), 0.0, stream(X)), "lob_matmul_nn")
    return Y.reshape(*batch_shape, M, C)


def tn_matmul(P: torch.Tensor, Q: torch.Tensor, out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """Out = P^T Q, reduction over the long (row) dimension.  P (*bp, N, I), Q (*bq, N, J)."""
    require_cuda(P, Q)
    lib = _lib.load()
    N, I = P.shape[-2:]
    J = Q.shape[-1]
    if Q.shape[-2] != N:
        raise RuntimeError(f"Size mismatch: {tuple(P.shape)}^T @ {tuple(Q.shape)}")
    batch_shape = torch.broadcast_shapes(P.shape[:-2], Q.shape[:-2])
    B = _numel(batch_shape)
    if _numel(P.shape[:-2]) == 1:
        Pf, p_bs = P.reshape(1, N, I).contiguous(), 0
    else:
        Pf, p_bs = _flat3(P.expand(*batch_shape, N, I)), N * I
    if _numel(Q.shape[:-2]) == 1:
        Qf, q_bs = Q.reshape(1, N, J).contiguous(), 0
    else:
        Qf, q_bs = _flat3(Q.expand(*batch_shape, N, J)), N * J
    out_dtype = out_dtype or P.dtype
    out = torch.empty(B, I, J, dtype=out_dtype, device=P.device)
    ws = workspace(lib.lob_tn_matmul_workspace_bytes(B, N, I, J), P.device)
    check(
        lib.lob_tn_matmul(dt(P), _lib._DT[out_dtype], B, N, I, J, ptr(Pf), p_bs, ptr(Qf), q_bs, ptr(out), ptr(ws),
                          stream(P)),
        "lob_tn_matmul",
    )
    return out.reshape(*batch_shape, I, J)


def scale_rows(X: torch.Tensor, d: torch.Tensor, mode: str) -> torch.Tensor:
    """rows of X (*b, N, C) scaled by f(d (*b, N) or (*b, 1)); mode in {"mul","div","mul_sqrt","div_sqrt"}"""
    require_cuda(X, d)
    lib = _lib.load()
    modes = {"mul": 0, "div": 1, "mul_sqrt": 2, "div_sqrt": 3}
    batch_shape = torch.broadcast_shapes(X.shape[:-2], d.shape[:-1])
    N, C = X.shape[-2:]
    Xf = _flat3(X.expand(*batch_shape, N, C))
    dd, d_bs, d_st = _diag_args(d, batch_shape, N)
    out = torch.empty_like(Xf)
    check(lib.lob_scale_rows(dt(X), Xf.shape[0], N, C, ptr(Xf), ptr(dd), d_bs, d_st, modes[mode], ptr(out), stream(X)),
          "lob_scale_rows")
    return out.reshape(*batch_shape, N, C)


# ------------------------------------------------------------------------------------------------------------
# pivoted Cholesky + preconditioner
# ------------------------------------------------------------------------------------------------------------
def _pivchol_finish(lib, Lt, perm, m_out, batch_shape, N, rankmax):
    m = int(m_out.item())  # the factor's width is data dependent: one host read, like the reference's per-step sync
    B = Lt.shape[0]
    L = torch.empty(B, N, m, dtype=Lt.dtype, device=Lt.device)
    check(lib.lob_transpose_rows(dt(Lt), B, rankmax, N, m, ptr(Lt), ptr(L), stream(Lt)), "lob_transpose_rows")
    return L.reshape(*batch_shape, N, m), perm.reshape(*batch_shape, N)


def pivoted_cholesky_dense(A: torch.Tensor, rank: int, tol: float) -> Tuple[torch.Tensor, torch.Tensor]:
    require_cuda(A)
    lib = _lib.load()
    batch_shape = A.shape[:-2]
    N = A.shape[-1]
    Af = _flat3(A)
    B = Af.shape[0]
    rankmax = min(int(rank), N)
    Lt = torch.zeros(B, rankmax, N, dtype=A.dtype, device=A.device)
    perm = torch.empty(B, N, dtype=torch.int64, device=A.device)
    m_out = torch.zeros(1, dtype=torch.int32, device=A.device)
    ws = workspace(lib.lob_pivchol_workspace_bytes(B, N, rankmax), A.device)
    check(
        lib.lob_pivchol_dense(dt(A), B, N, rankmax, float(tol), ptr(Af), N, N * N, ptr(Lt), ptr(perm), ptr(m_out),
                              ptr(ws), stream(A)),
        "lob_pivchol_dense",
    )
    return _pivchol_finish(lib, Lt, perm, m_out, batch_shape, N, rankmax)


def pivoted_cholesky_kron(factors: Sequence[torch.Tensor], batch_shape, rank: int, tol: float):
    require_cuda(*factors)
    lib = _lib.load()
    if not 1 <= len(factors) <= 4:
        raise _lib.LobError("Kronecker pivoted Cholesky supports 1..4 factors")
    B = _numel(batch_shape)
    sizes = [int(f.shape[-1]) for f in factors]
    N = int(math.prod(sizes))
    flat, strides = [], []
    for f in factors:
        if _numel(f.shape[:-2]) == 1:
            flat.append(f.reshape(1, *f.shape[-2:]).contiguous())
            strides.append(0)
        else:
            ff = _flat3(f.expand(*batch_shape, *f.shape[-2:]))
            flat.append(ff)
            strides.append(ff.shape[-1] * ff.shape[-2])
    ref = flat[0]
    rankmax = min(int(rank), N)
    Lt = torch.zeros(B, rankmax, N, dtype=ref.dtype, device=ref.device)
    perm = torch.empty(B, N, dtype=torch.int64, device=ref.device)
    m_out = torch.zeros(1, dtype=torch.int32, device=ref.device)
    ws = workspace(lib.lob_pivchol_workspace_bytes(B, N, rankmax), ref.device)
    nf = len(flat)
    c_sizes = (ctypes.c_int64 * nf)(*sizes)
    c_ptrs = (ctypes.c_void_p * nf)(*[f.data_ptr() for f in flat])
    c_strides = (ctypes.c_int64 * nf)(*strides)
    check(
        lib.lob_pivchol_kron(dt(ref), B, nf, c_sizes, c_ptrs, c_strides, rankmax, float(tol), ptr(Lt), ptr(perm),
                             ptr(m_out), ptr(ws), stream(ref)),
        "lob_pivchol_kron",
    )
    return _pivchol_finish(lib, Lt, perm, m_out, batch_shape, N, rankmax)


def pivoted_cholesky_toeplitz(col: torch.Tensor, rank: int, tol: float):
    require_cuda(col)
    lib = _lib.load()
    batch_shape = col.shape[:-1]
    N = col.shape[-1]
    cf = col.contiguous().reshape(-1, N)
    B = cf.shape[0]
    rankmax = min(int(rank), N)
    Lt = torch.zeros(B, rankmax, N, dtype=col.dtype, device=col.device)
    perm = torch.empty(B, N, dtype=torch.int64, device=col.device)
    m_out = torch.zeros(1, dtype=torch.int32, device=col.device)
    ws = workspace(lib.lob_pivchol_workspace_bytes(B, N, rankmax), col.device)
    check(
        lib.lob_pivchol_toeplitz(dt(col), B, N, ptr(cf), N, rankmax, float(tol), ptr(Lt), ptr(perm), ptr(m_out),
                                 ptr(ws), stream(col)),
        "lob_pivchol_toeplitz",
    )
    return _pivchol_finish(lib, Lt, perm, m_out, batch_shape, N, rankmax)


def pivoted_cholesky_rows(op, rank: int, tol: float, poll_every: int = 8):
    """Pivoted Cholesky of ANY operator that implements ``_get_indices`` and ``_approx_diagonal`` (Root, Sum, user
    classes): the reference's generic route (functions/_pivoted_cholesky.py:57-98 -> utils/permutation.py:76-87 ->
    ``LinearOperator._get_indices``), with the pivot search / row update / stop rule in the same device kernels as the
    dense case.  The pivot indices stay on the device; the host polls the stop flag every ``poll_every`` steps only to
    skip useless row fetches after an early stop."""
    lib = _lib.load()
    batch_shape = op.batch_shape
    N = op.size(-1)
    B = _numel(batch_shape)
    diag = op._approx_diagonal()
    require_cuda(diag)
    dev, dty = diag.device, diag.dtype
    diag = diag.expand(*batch_shape, N).reshape(B, N).contiguous()
    rankmax = min(int(rank), N)
    Lt = torch.zeros(B, rankmax, N, dtype=dty, device=dev)
    perm = torch.empty(B, N, dtype=torch.int64, device=dev)
    pi = torch.zeros(B, dtype=torch.int64, device=dev)
    status = torch.zeros(2, dtype=torch.int32, device=dev)
    ws = workspace(lib.lob_pivchol_workspace_bytes(B, N, rankmax), dev)
    st = stream(diag)
    check(lib.lob_pivchol_rows_begin(dt(diag), B, N, rankmax, ptr(diag), ptr(perm), ptr(ws), st), "lob_pivchol_rows_begin")
    # index tensors of the row gather  op[(*batch_indices, pi_m, arange(N))]  (utils/permutation.py:66-87)
    cols = torch.arange(N, device=dev).expand(*batch_shape, N)
    batch_idx = []
    for i, sz in enumerate(batch_shape):
        shape = [1] * (len(batch_shape) + 1)
        shape[i] = sz
        batch_idx.append(torch.arange(sz, device=dev).reshape(shape).expand(*batch_shape, N))
    for m in range(rankmax):
        check(lib.lob_pivchol_rows_pivot(dt(diag), B, N, rankmax, m, float(tol), ptr(Lt), ptr(perm), ptr(pi), ptr(ws), st),
              "lob_pivchol_rows_pivot")
        if m + 1 < N:
            rows_idx = pi.reshape(*batch_shape, 1).expand(*batch_shape, N)
            row = op._get_indices(rows_idx, cols, *batch_idx)
            row = row.to(dty).expand(*batch_shape, N).reshape(B, N).contiguous()
            check(lib.lob_pivchol_rows_update(dt(diag), B, N, rankmax, m, ptr(row), ptr(Lt), ptr(ws), st),
                  "lob_pivchol_rows_update")
        if poll_every and (m + 1) % poll_every == 0 and m + 1 < rankmax:
            check(lib.lob_pivchol_rows_status(dt(diag), B, N, rankmax, ptr(status), ptr(status[1:]), ptr(ws), st),
                  "lob_pivchol_rows_status")
            if int(status[1].item()) == 0:
                break
    check(lib.lob_pivchol_rows_status(dt(diag), B, N, rankmax, ptr(status), ptr(status[1:]), ptr(ws), st),
          "lob_pivchol_rows_status")
    return _pivchol_finish(lib, Lt, perm, status[:1], batch_shape, N, rankmax)


class AddedDiagPreconditioner:
    """M = L L^T + D from a pivoted-Cholesky factor (added_diag_linear_operator.py:144-184).

    Holds Q (*b, N, k) with M^-1 v = (v - Q Q^T v)/s (constant diagonal) or v/d - Q Q^T v (general diagonal) and
    logdet(M).  Built through the k x k Gram matrix in double (see csrc/precond.cu)."""

    def __init__(self, L: torch.Tensor, diag: torch.Tensor, constant: bool):
        require_cuda(L, diag)
        lib = _lib.load()
        self.batch_shape = L.shape[:-2]
        N, k = L.shape[-2:]
        self.N, self.k, self.constant = N, k, constant
        B = _numel(self.batch_shape)
        dev, dty = L.device, L.dtype
        if constant:
            self.noise = diag[..., :1].expand(*self.batch_shape, 1).reshape(B, 1).contiguous()  # sigma^2 per batch elt
            Ls = L
        else:
            self.noise = diag.expand(*self.batch_shape, N).reshape(B, N).contiguous()
            Ls = scale_rows(L, self.noise.reshape(*self.batch_shape, N), "div_sqrt")  # D^-1/2 L (:176-178)
        # Gram matrix L^T L (k x k, contraction over N) in double.  fp32 factors: tensor cores with split-K so that no
        # accumulator sees more than 512 contraction indices (the truncating fp32 accumulate would otherwise bias the
        # all-positive diagonal sums by ~1.5e-8 per index), partial sums added in double; fp64 factors: CUDA cores.
        G = None
        if dty == torch.float32 and N >= 2048:
            G = gemm3x(Ls.reshape(B, N, k), Ls.reshape(B, N, k), trans_a=True, splits=max(2, -(-N // 512)),
                       out_dtype=torch.float64)
        if G is None:
            G = tn_matmul(Ls, Ls, out_dtype=torch.float64)
        G = G.reshape(B, k, k)
        rinv = torch.empty(B, k, k, dtype=dty, device=dev)
        logdet_r = torch.empty(B, dtype=dty, device=dev)
        info = torch.zeros(B, dtype=torch.int32, device=dev)
        ws = workspace(B * k * k * 8, dev)
        check(
            lib.lob_precond_factor(dt(L), B, k, ptr(G), 1.0, ptr(self.noise) if constant else None, 1, ptr(rinv),
                                   ptr(logdet_r), ptr(info), ptr(ws), stream(L)),
            "lob_precond_factor",
        )
        # fp32 ranks that are not a multiple of 4 (the reference's default is 15): Q gets zero columns up to the next
        # multiple of 4 (at least 8), so that its rows are 16-byte addressable and both products of the apply take
        # the TMA / cp.async kernels instead of the generic CUDA-core ones; zero columns change no product.
        kp = max(8, -(-k // 4) * 4) if dty == torch.float32 else k
        if kp != k:
            rinv = torch.cat([rinv, rinv.new_zeros(B, k, kp - k)], dim=-1)
        Q = gemm3x(Ls.reshape(B, N, k), rinv) if dty == torch.float32 and N >= 2048 else None  # Q1 = L R^-1
        if Q is None:
            Q = matmul_nn(Ls.reshape(B, N, k), rinv)
        if constant:
            self.Q = Q
            # logdet M = 2 sum log|R_ii| + (N - k) log s   (:170-172); tiny (B,) op on the control path
            self.logdet = logdet_r + (N - k) * self.noise[:, 0].log()
            self._alpha = -self.noise[:, 0].reciprocal()  # z = (r - w)/s = alpha * w + (1/s) r
            self._dscale = self.noise.reciprocal()  # (B, 1)
        else:
            self.Q = scale_rows(Q, self.noise, "div_sqrt")  # D^-1/2 Q1 (:179)
            self.logdet = logdet_r + self.noise.log().sum(-1)  # :182-183
            self._alpha = torch.full((B,), -1.0, dtype=dty, device=dev)  # z = r/d - w
            self._dscale = self.noise.reciprocal()  # (B, N)
        self.logdet = self.logdet.reshape(self.batch_shape) if len(self.batch_shape) else self.logdet.squeeze()
        self.info = info  # (B,) int32, != 0: non-positive pivot in the k x k factorisation

    def _apply(self, v: torch.Tensor, want_dots: bool):
        require_cuda(v)
        batch_shape = torch.broadcast_shapes(self.batch_shape, v.shape[:-2])
        N, C = v.shape[-2:]
        vf = _flat3(v.expand(*batch_shape, N, C))
        B = vf.shape[0]
        if B != self.Q.shape[0]:
            raise _lib.LobError("preconditioner batch shape and right-hand-side batch shape differ")
        # Q^T v: long (K = N) contraction on the CUDA cores with round-to-nearest fp32 accumulation.  The tensor-core
        # kernel truncates when it adds into its accumulator, a bias ~3e-9*N that is harmless in the operator matmul
        # but, applied to the preconditioner, shifts logdet by O(k * cond * bias) (measured 7e-5 relative at N=5000).
        t = tn_matmul(self.Q, vf)  # (B, k, C)
        out = dense_matmul(self.Q, t, d=self._dscale, want_dots=want_dots, E=vf, alpha=self._alpha)
        if want_dots:
            z, dots, n_parts = out
            return z.reshape(*batch_shape, N, C), dots, n_parts
        return out.reshape(*batch_shape, N, C)

    def __call__(self, v: torch.Tensor) -> torch.Tensor:
        """precondition_closure (added_diag_linear_operator.py:135-140): z = (v - Q Q^T v)/s resp. v/d - Q Q^T v."""
        squeeze = v.dim() == 1
        if squeeze:
            v = v.unsqueeze(-1)
        z = self._apply(v, False)
        return z.squeeze(-1) if squeeze else z

    def fused(self, v: torch.Tensor):
        """(z, partial sums of <v, z>, n_parts): the preconditioned residual and linear_cg's <r, z> in one pass."""
        return self._apply(v, True)


# ------------------------------------------------------------------------------------------------------------
# probes, column reductions, quadrature
# ------------------------------------------------------------------------------------------------------------
def probe_assemble(z_root: Optional[torch.Tensor], eps_diag: torch.Tensor, d: Optional[torch.Tensor]):
    """probes (*b, N, S), norms (*b, 1, S) from z_root (*b, N, S) and eps_diag (S, *b, N)
    (functions/_inv_quad_logdet.py:107-110)."""
    require_cuda(z_root, eps_diag, d)
    lib = _lib.load()
    S = eps_diag.shape[0]
    batch_shape = eps_diag.shape[1:-1]
    N = eps_diag.shape[-1]
    B = _numel(batch_shape)
    ef = eps_diag.contiguous().reshape(S, B, N)
    zf = None if z_root is None else _flat3(z_root)
    dd, d_bs, d_st = _diag_args(d, batch_shape, N)
    probes = torch.empty(B, N, S, dtype=ef.dtype, device=ef.device)
    norms = torch.empty(B, 1, S, dtype=ef.dtype, device=ef.device)
    ws = workspace(lib.lob_colred_workspace_bytes(B, N, S), ef.device)
    check(
        lib.lob_probe_assemble(dt(ef), B, N, S, ptr(zf), ptr(ef), ptr(dd), d_bs, d_st, ptr(probes), ptr(norms), ptr(ws),
                               stream(ef)),
        "lob_probe_assemble",
    )
    return probes.reshape(*batch_shape, N, S), norms.reshape(*batch_shape, 1, S)


def col_dots(U: torch.Tensor, u_off: int, V: torch.Tensor, v_off: int, R: int) -> torch.Tensor:
    """out (*b, R) = sum_n U[..., n, u_off + j] V[..., n, v_off + j]"""
    require_cuda(U, V)
    lib = _lib.load()
    batch_shape = U.shape[:-2]
    Uf, Vf = _flat3(U), _flat3(V)
    B, N, Cu = Uf.shape
    Cv = Vf.shape[-1]
    out = torch.empty(B, R, dtype=U.dtype, device=U.device)
    ws = workspace(lib.lob_colred_workspace_bytes(B, N, R), U.device)
    check(lib.lob_col_dots(dt(U), B, N, R, ptr(Uf), Cu, u_off, ptr(Vf), Cv, v_off, ptr(out), ptr(ws), stream(U)),
          "lob_col_dots")
    return out.reshape(*batch_shape, R)


def tridiag_eigh_slq(t_mat: torch.Tensor, n: int, want_evals=False, want_evecs=False, want_logdet=True):
    """t_mat (*lead, T, T) symmetric tridiagonal.  evals (*lead, T) / evecs (*lead, T, T) accept any leading shape like
    the reference's eigh-based version (utils/lanczos.py:167-189; ``lanczos_tridiag`` squeezes the probe dimension for a
    single vector, so a bare (T, T) matrix arrives here too).  The quadrature ``logdet`` (*b) needs the
    (S, *b, T, T) layout: it sums over the leading probe dimension."""
    require_cuda(t_mat)
    lib = _lib.load()
    if t_mat.dim() < 2 or t_mat.shape[-1] != t_mat.shape[-2]:
        raise RuntimeError(f"expected (*, T, T) tridiagonal matrices, got {tuple(t_mat.shape)}")
    if want_logdet and t_mat.dim() < 3:
        raise RuntimeError("the quadrature needs t_mat of shape (num_probes, *batch, T, T)")
    lead = t_mat.shape[:-2]
    T = t_mat.shape[-1]
    if want_logdet:
        S, batch_shape = int(lead[0]), lead[1:]
        B = _numel(batch_shape)
    else:  # eigendecomposition only: every matrix is its own problem
        S, batch_shape, B = _numel(lead), lead, 1
    tf = t_mat.contiguous()
    dev, dty = t_mat.device, t_mat.dtype
    evals = torch.empty(S, B, T, dtype=dty, device=dev) if want_evals else None
    evecs = torch.empty(S, B, T, T, dtype=dty, device=dev) if want_evecs else None
    logdet = torch.empty(B, dtype=dty, device=dev) if want_logdet else None
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    ws = workspace(lib.lob_tridiag_workspace_bytes(S, B, T, 1 if want_evecs else 0), dev)
    check(
        lib.lob_tridiag_eigh_slq(dt(t_mat), S, B, T, int(n), ptr(tf), ptr(evals), ptr(evecs), ptr(logdet), ptr(info),
                                 ptr(ws), stream(t_mat)),
        "lob_tridiag_eigh_slq",
    )
    out = {}
    if want_evals:
        out["evals"] = evals.reshape(*lead, T)
    if want_evecs:
        out["evecs"] = evecs.reshape(*lead, T, T)
    if want_logdet:
        out["logdet"] = logdet.reshape(batch_shape)
    return out


# ------------------------------------------------------------------------------------------------------------
# structured matmuls
# ------------------------------------------------------------------------------------------------------------
def _kron_tensor_core_ok(factors, X) -> bool:
    """The tcgen05 chain needs TMA-addressable views: fp32, square factors with sizes that are multiples of 4 (leading
    dimensions of the factor and of every mode's (n_i x post) slab)."""
    if X.dtype != torch.float32 or len(factors) < 2:
        return False
    sizes = [int(f.shape[-1]) for f in factors]
    if any(f.shape[-2] != f.shape[-1] for f in factors) or any(n % 4 for n in sizes):
        return False
    return int(math.prod(sizes)) == X.shape[-2]


def _kron_matmul_tc(factors, X, d, want_dots=False):
    """(K1 (x) ... (x) Km) X (+ d (.) X) as a chain of tensor-core GEMMs (csrc/gemm3x.cu) on column planes.

    The vector block X (B, N, C) has C = 33 columns: a slab with a 33-element leading dimension is not TMA-addressable,
    so the chain works on the transposed planes Xt (B, C, n1, ..., nm) (one transposing pass in, one out -- the same
    kernels as the Toeplitz pad / unpad, with the + d (.) X of AddedDiag fused into the way out).  With pre = n1..n_{i-1}
    and post = n_{i+1}..n_m, mode i is the batched GEMM  Y[b,c,p] (n_i x post) = K_i[b] X[b,c,p] (n_i x post) for i < m
    and, for the last mode,  Y[b,c] (pre x n_m) = X[b,c] (pre x n_m) K_m[b]^T.  The index order never changes, so there
    are no transposing copies between the modes (kronecker_product_linear_operator.py:34-45 has one per factor):
    m + 2 passes over the vector in total, every product on the tensor cores."""
    lib = _lib.load()
    batch_shape = torch.broadcast_shapes(X.shape[:-2], *[f.shape[:-2] for f in factors])
    B = _numel(batch_shape)
    N, C = X.shape[-2:]
    Xf = _flat3(X.expand(*batch_shape, N, C))
    fl = []
    for f in factors:
        n = f.shape[-1]
        fl.append(f.reshape(1, n, n).contiguous() if _numel(f.shape[:-2]) == 1 else _flat3(f.expand(*batch_shape, n, n)))
    sizes = [int(f.shape[-1]) for f in factors]
    dd, d_bs, d_st = _diag_args(d, batch_shape, N)
    Y = torch.empty(B, N, C, dtype=X.dtype, device=X.device)
    # linear_cg's <p, A p> out of the chain's last pass (it holds both X and Y): one partial sum per row block, column
    n_parts = int(lib.lob_toeplitz_unpad_parts(dt(X), N, C)) if want_dots else 0
    dots = torch.empty(B, n_parts, C, dtype=torch.float64, device=X.device) if n_parts else None
    # batch chunks keep every mode's GEMM batch (B * C * pre) inside the launch limit
    max_pre = int(math.prod(sizes[:-2])) if len(sizes) > 2 else 1
    bchunk = max(1, min(B, 65535 // (C * max_pre)))
    for b0 in range(0, B, bchunk):
        b1 = min(B, b0 + bchunk)
        nb = b1 - b0
        cur = torch.empty(nb, C, N, dtype=X.dtype, device=X.device)
        check(lib.lob_toeplitz_pad(dt(X), nb, N, C, N, ptr(Xf[b0:b1]), ptr(cur), stream(X)), "lob_toeplitz_pad")
        pre = 1
        for i, n in enumerate(sizes):
            post = N // (pre * n)
            Ki = fl[i] if fl[i].shape[0] == 1 else fl[i][b0:b1]
            shared = Ki.shape[0] == 1
            if i < len(sizes) - 1:
                nbatch = nb * C * pre
                out = gemm3x(Ki, cur.reshape(nbatch, n, post), a_div=(nbatch if shared else C * pre), batch=nbatch)
            else:
                nbatch = nb * C
                out = gemm3x(cur.reshape(nbatch, pre, n), Ki, trans_b=True, b_div=(nbatch if shared else C),
                             batch=nbatch)
            if out is None:
                raise _lib.LobError("Kronecker tensor-core chain: operand not TMA-addressable")
            cur = out
            pre *= n
        dd_c = dd if (dd is None or d_bs == 0) else dd[b0:b1]
        check(
            lib.lob_toeplitz_unpad(dt(X), nb, N, C, N, ptr(cur), 1.0, ptr(Xf[b0:b1]), ptr(dd_c), d_bs, d_st,
                                   ptr(Y[b0:b1]), ptr(dots[b0:b1]) if dots is not None else None, stream(X)),
            "lob_toeplitz_unpad",
        )
    Y = Y.reshape(*batch_shape, N, C)
    return (Y, dots, n_parts) if want_dots else Y


def kron_matmul(factors: Sequence[torch.Tensor], X: torch.Tensor, d: Optional[torch.Tensor] = None,
                want_dots: bool = False):
    """(K1 (x) K2 (x) ...) X (+ d (.) X)  (kronecker_product_linear_operator.py:34-45; added_diag_linear_operator.py
    :72-76 for the fused diagonal).  fp32 operators with TMA-addressable factors run as a tensor-core GEMM chain
    (``_kron_matmul_tc``); everything else (fp64, odd factor sizes) as one fused CUDA-core mode product per factor."""
    require_cuda(X, d, *factors)
    lib = _lib.load()
    if _kron_tensor_core_ok(factors, X):
        return _kron_matmul_tc(factors, X, d, want_dots)
    batch_shape = torch.broadcast_shapes(X.shape[:-2], *[f.shape[:-2] for f in factors])
    B = _numel(batch_shape)
    Ntot, C = X.shape[-2:]
    cur = _flat3(X.expand(*batch_shape, Ntot, C))
    for f in factors:
        n = f.shape[-1]
        if f.shape[-2] != n:
            raise _lib.LobError("Kronecker matmul kernel expects square factors")
        Q = Ntot // n
        if _numel(f.shape[:-2]) == 1:
            ff, k_bs = f.reshape(1, n, n).contiguous(), 0
        else:
            ff, k_bs = _flat3(f.expand(*batch_shape, n, n)), n * n
        out = torch.empty_like(cur)
        check(lib.lob_kron_mode_matmul(dt(X), B, n, Q, C, ptr(ff), k_bs, ptr(cur), ptr(out), stream(X)),
              "lob_kron_mode_matmul")
        cur = out
    cur = cur.reshape(*batch_shape, Ntot, C)
    if d is not None:
        cur = cur.add_(scale_rows(X.expand(*batch_shape, Ntot, C), d, "mul"))
    return (cur, None, 0) if want_dots else cur  # no fused <X, Y> on the CUDA-core mode kernels: linear_cg adds its own


def _next_pow2(n: int) -> int:
    return 1 << (int(n) - 1).bit_length()


def toeplitz_embed_fft(col: torch.Tensor):
    """FFT of the circulant embedding of a symmetric Toeplitz column: returns (fc (B, L/2+1) complex, L)."""
    require_cuda(col)
    lib = _lib.load()
    N = col.shape[-1]
    L = _next_pow2(2 * N)
    cf = col.contiguous().reshape(-1, N)
    B = cf.shape[0]
    c = torch.empty(B, L, dtype=col.dtype, device=col.device)
    check(lib.lob_toeplitz_embed(dt(col), B, N, L, ptr(cf), N, ptr(c), stream(col)), "lob_toeplitz_embed")
    return torch.fft.rfft(c), L  # cuFFT R2C


TOEPLITZ_SCRATCH_BYTES = 4.5 * 2**30  # per (B, C, L) scratch array of one batch chunk of toeplitz_matmul


def toeplitz_matmul(col: torch.Tensor, X: torch.Tensor, d: Optional[torch.Tensor] = None, fc_cache=None,
                    want_dots: bool = False):
    """Symmetric Toeplitz matmul through a length-L (power of two >= 2N) real circulant embedding
    (utils/toeplitz.py:131-149 uses length 2N-1 complex FFTs; any L >= 2N-1 gives the same product).
    Optionally fuses + d (.) X; ``want_dots``: returns ``(Y, dots, n_parts)`` with the (B, n_parts, C) partial sums of
    X * Y out of the unpack pass (linear_cg's <p, A p>)."""
    require_cuda(col, X, d)
    lib = _lib.load()
    N, C = X.shape[-2:]
    batch_shape = torch.broadcast_shapes(col.shape[:-1], X.shape[:-2])
    B = _numel(batch_shape)
    Xf = _flat3(X.expand(*batch_shape, N, C))
    if fc_cache is None:
        fc_cache = toeplitz_embed_fft(col)
    fc, L = fc_cache
    if fc.shape[0] == 1:
        fc_bs = 0
    elif fc.shape[0] == B:
        fc_bs = fc.shape[-1]
    else:  # the column's batch is a proper sub-batch of the broadcast batch: materialise the broadcast spectrum
        fc = fc.reshape(*col.shape[:-1], fc.shape[-1]).expand(*batch_shape, fc.shape[-1]).reshape(B, -1).contiguous()
        fc_bs = fc.shape[-1]
    Y = torch.empty(B, N, C, dtype=X.dtype, device=X.device)
    dd, d_bs, d_st = _diag_args(d, batch_shape, N)
    # The embedding of a symmetric Toeplitz matrix has a real spectrum: column PAIRS ride one complex transform
    # (csrc/structured.cu, "complex FFTs of column pairs").  Scratch: two (B, ceil(C/2), L) complex arrays (transform
    # input and output) -- 2 x 17.8 GB at BASELINE config 4 (B = 64, N = 2^20, 33 columns) -- so the product runs in
    # batch chunks whose arrays stay below ~4.5 GB each (16 elements at config 4); batch elements are independent.
    rows_parts = int(lib.lob_toeplitz_unpack_parts(dt(X), N, C))
    if rows_parts == 0:
        # more columns than one shared-memory tile of the pack / unpack kernels holds: column blocks of 64
        outs = [toeplitz_matmul(col, X[..., c0:c0 + 64].contiguous(), d, fc_cache=(fc, L)) for c0 in range(0, C, 64)]
        Yc = torch.cat(outs, dim=-1)
        return (Yc, None, 0) if want_dots else Yc
    n_parts = rows_parts if want_dots else 0
    dots = torch.empty(B, n_parts, C, dtype=torch.float64, device=X.device) if n_parts else None
    fr = fc.real.contiguous()  # (B | 1, L / 2 + 1)
    fr_bs = 0 if fc_bs == 0 else fr.shape[-1]
    P = (C + 1) // 2
    cdtype = torch.complex64 if X.dtype == torch.float32 else torch.complex128
    bits_dtype = torch.int32 if X.dtype == torch.float32 else torch.int64
    chunk = max(1, min(B, int(TOEPLITZ_SCRATCH_BYTES // (2 * P * L * X.element_size()))))
    for b0 in range(0, B, chunk):
        b1 = min(B, b0 + chunk)
        nb = b1 - b0
        xs = Xf[b0:b1]
        maxbits = torch.empty(nb, C, dtype=bits_dtype, device=X.device)
        check(lib.lob_toeplitz_colmax(dt(X), nb, N, C, ptr(xs), ptr(maxbits), stream(X)), "lob_toeplitz_colmax")
        zt = torch.empty(nb, P, L, dtype=cdtype, device=X.device)
        check(lib.lob_toeplitz_pack(dt(X), nb, N, C, L, ptr(xs), ptr(maxbits), ptr(zt), stream(X)), "lob_toeplitz_pack")
        # cuFFT C2C, batched over (B, P); out-of-place on purpose: with out= torch transforms into a temporary and
        # then copies (a full extra pass, measured)
        zt = torch.fft.fft(zt, dim=-1)
        fr_c = fr if fr_bs == 0 else fr[b0:b1]
        check(lib.lob_toeplitz_mulr(dt(X), nb, P, L, ptr(fr_c), fr_bs, ptr(zt), stream(X)), "lob_toeplitz_mulr")
        # un-normalised inverse (norm="forward" puts the 1/L on the forward transform, which we did not ask for): the
        # 1/L goes into the unpack kernel instead of a separate full pass
        zt = torch.fft.ifft(zt, dim=-1, norm="forward")
        dd_c = dd if (dd is None or d_bs == 0) else dd[b0:b1]
        check(
            lib.lob_toeplitz_unpack(dt(X), nb, N, C, L, ptr(zt), 1.0 / L, ptr(maxbits), ptr(xs), ptr(dd_c), d_bs, d_st,
                                    ptr(Y[b0:b1]), ptr(dots[b0:b1]) if dots is not None else None, stream(X)),
            "lob_toeplitz_unpack",
        )
        del zt
    Y = Y.reshape(*batch_shape, N, C)
    return (Y, dots, n_parts) if want_dots else Y


def cap_solve(G: torch.Tensor, W: torch.Tensor):
    """W <- (I + G)^-1 W, logdet(I + G).  G (*bg, k, k) float64, W (*b, k, C)."""
    require_cuda(G, W)
    lib = _lib.load()
    batch_shape = W.shape[:-2]
    k, C = W.shape[-2:]
    Wf = _flat3(W).clone()
    B = Wf.shape[0]
    if _numel(G.shape[:-2]) == 1:
        Gf, g_bs = G.reshape(1, k, k).contiguous(), 0
    else:
        Gf, g_bs = _flat3(G.expand(*batch_shape, k, k)), k * k
    logdet = torch.empty(B, dtype=W.dtype, device=W.device)
    info = torch.zeros(B, dtype=torch.int32, device=W.device)
    ws = workspace(lib.lob_cap_solve_workspace_bytes(B, k, C), W.device)
    check(lib.lob_cap_solve(dt(W), B, k, C, ptr(Gf), g_bs, ptr(Wf), ptr(logdet), ptr(info), ptr(ws), stream(W)),
          "lob_cap_solve")
    return Wf.reshape(*batch_shape, k, C), logdet.reshape(batch_shape), info


# ------------------------------------------------------------------------------------------------------------
# Lanczos with full re-orthogonalisation (utils/lanczos.py:9-164)
# ------------------------------------------------------------------------------------------------------------
def lanczos_init(init_vecs: torch.Tensor, q0: torch.Tensor):
    """q0 (B, N, C) <- init_vecs / ||init_vecs||_2 column-wise."""
    require_cuda(init_vecs, q0)
    lib = _lib.load()
    B, N, C = init_vecs.shape
    check(lib.lob_lanczos_init(dt(init_vecs), B, N, C, ptr(init_vecs), ptr(q0), stream(init_vecs)), "lob_lanczos_init")


def lanczos_step(mode: int, k: int, w: Optional[torch.Tensor], q_mat: torch.Tensor, t_mat: torch.Tensor,
                 flags: torch.Tensor, tol: float):
    """One fused Lanczos iteration on q_mat (T, B, N, C), t_mat (T, T, B, C); flags = int32[2] decision words."""
    require_cuda(q_mat, t_mat, flags, w)
    lib = _lib.load()
    T, B, N, C = q_mat.shape
    check(
        lib.lob_lanczos_step(dt(q_mat), mode, B, N, C, T, k, ptr(w), ptr(q_mat), ptr(t_mat), ptr(flags), float(tol),
                             stream(q_mat)),
        "lob_lanczos_step",
    )


# ------------------------------------------------------------------------------------------------------------
# backward pass
# ------------------------------------------------------------------------------------------------------------
def _col_weights(w: Optional[torch.Tensor], batch_shape, C: int, like: torch.Tensor):
    if w is None:
        return None
    return w.to(like.dtype).expand(*batch_shape, C).reshape(-1, C).contiguous()


def bilinear_dense(left: torch.Tensor, right: torch.Tensor, w: Optional[torch.Tensor] = None,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """sum_c w[..., c] left[..., :, c] right[..., :, c]^T  -> (*b, N, M): DenseLinearOperator._bilinear_derivative
    (dense_linear_operator.py:69-71) with the callers' column scalings folded in.  ``out``: accumulate into it."""
    require_cuda(left, right, w, out)
    lib = _lib.load()
    batch_shape = torch.broadcast_shapes(left.shape[:-2], right.shape[:-2])
    N, C = left.shape[-2:]
    M = right.shape[-2]
    if right.shape[-1] != C:
        raise RuntimeError(f"Size mismatch: {tuple(left.shape)} vs {tuple(right.shape)}")
    Lf = _flat3(left.expand(*batch_shape, N, C))
    Rf = _flat3(right.expand(*batch_shape, M, C))
    B = Lf.shape[0]
    wf = _col_weights(w, batch_shape, C, left)
    if out is None and left.dtype == torch.float32 and N >= 1024 and M >= 1024 and M % 4 == 0:
        # tensor cores: G = L diag(w) R^T is a GEMM with K = C.  The 33-column blocks are not TMA-addressable (16-byte
        # rows), so both factors are copied once into 4-column-aligned buffers (1 % of the gradient's bytes); the
        # B N^2 result is written by the TMA epilogue of csrc/gemm3x.cu.
        Cp = -(-C // 4) * 4
        Lp = torch.zeros(B, N, Cp, dtype=left.dtype, device=left.device)
        Rp = torch.zeros(B, M, Cp, dtype=left.dtype, device=left.device)
        Lp[..., :C] = Lf if wf is None else Lf * wf.unsqueeze(-2)
        Rp[..., :C] = Rf
        G = gemm3x(Lp, Rp, trans_b=True)
        if G is not None:
            return G.reshape(*batch_shape, N, M)
    acc = 1
    if out is None:
        out = torch.empty(B, N, M, dtype=left.dtype, device=left.device)
        acc = 0
    elif not out.is_contiguous() or out.numel() != B * N * M:
        raise _lib.LobError("bilinear_dense: `out` must be a contiguous (*batch, N, M) tensor")
    check(lib.lob_bilinear_dense(dt(left), B, N, M, C, ptr(Lf), ptr(Rf), ptr(wf), ptr(out), acc, stream(left)),
          "lob_bilinear_dense")
    return out.reshape(*batch_shape, N, M)


def bilinear_diag(left: torch.Tensor, right: torch.Tensor, w: Optional[torch.Tensor] = None) -> torch.Tensor:
    """sum_c w[..., c] left[..., n, c] right[..., n, c] -> (*b, N)  (diag_linear_operator.py:37-45)."""
    require_cuda(left, right, w)
    lib = _lib.load()
    batch_shape = torch.broadcast_shapes(left.shape[:-2], right.shape[:-2])
    N, C = left.shape[-2:]
    Lf = _flat3(left.expand(*batch_shape, N, C))
    Rf = _flat3(right.expand(*batch_shape, N, C))
    B = Lf.shape[0]
    wf = _col_weights(w, batch_shape, C, left)
    out = torch.empty(B, N, dtype=left.dtype, device=left.device)
    check(lib.lob_bilinear_diag(dt(left), B, N, C, ptr(Lf), ptr(Rf), ptr(wf), ptr(out), stream(left)),
          "lob_bilinear_diag")
    return out.reshape(*batch_shape, N)


def tri_inverse(Cm: torch.Tensor) -> torch.Tensor:
    """Inverse of (a batch of) lower-triangular k x k matrices."""
    require_cuda(Cm)
    lib = _lib.load()
    k = Cm.shape[-1]
    Cf = _flat3(Cm)
    out = torch.empty_like(Cf)
    check(lib.lob_tri_inverse(dt(Cm), Cf.shape[0], k, ptr(Cf), k, k * k, ptr(out), stream(Cm)), "lob_tri_inverse")
    return out.reshape(Cm.shape)


def toeplitz_bilinear_derivative(left: torch.Tensor, right: torch.Tensor, w: Optional[torch.Tensor] = None):
    """res[..., i] = sum_c w_c u_c^T (dT/dc_i) v_c, the gradient of a symmetric Toeplitz operator's column
    (utils/toeplitz.py:164-204; the reference runs two Toeplitz products per column): cross-correlations through a
    zero-padded real FFT of length L >= 2N, summed over the columns in the frequency domain."""
    require_cuda(left, right, w)
    lib = _lib.load()
    batch_shape = torch.broadcast_shapes(left.shape[:-2], right.shape[:-2])
    N, C = left.shape[-2:]
    L = _next_pow2(2 * N)
    Lf = _flat3(left.expand(*batch_shape, N, C))
    Rf = _flat3(right.expand(*batch_shape, N, C))
    B = Lf.shape[0]
    wf = _col_weights(w, batch_shape, C, left)
    out = torch.empty(B, N, dtype=left.dtype, device=left.device)
    chunk = max(1, min(B, int(TOEPLITZ_SCRATCH_BYTES // (C * L * left.element_size()))))
    for b0 in range(0, B, chunk):
        b1 = min(B, b0 + chunk)
        nb = b1 - b0
        spectra = []
        for X in (Lf, Rf):
            xt = torch.empty(nb, C, L, dtype=left.dtype, device=left.device)
            check(lib.lob_toeplitz_pad(dt(left), nb, N, C, L, ptr(X[b0:b1]), ptr(xt), stream(left)), "lob_toeplitz_pad")
            spectra.append(torch.fft.rfft(xt))
            del xt
        fu, fv = spectra
        H = fu.shape[-1]
        spec = torch.empty(nb, H, dtype=fu.dtype, device=left.device)
        check(lib.lob_toeplitz_cross_spectrum(dt(left), nb, C, H, ptr(fu), ptr(fv),
                                              ptr(None if wf is None else wf[b0:b1]), ptr(spec), stream(left)),
              "lob_toeplitz_cross_spectrum")
        del fu, fv, spectra
        y = torch.fft.irfft(spec, n=L, norm="forward")  # un-normalised C2R; the 1/L goes into the finishing kernel
        check(lib.lob_toeplitz_deriv_finish(dt(left), nb, N, L, ptr(y), 1.0 / L, ptr(out[b0:b1]), stream(left)),
              "lob_toeplitz_deriv_finish")
    return out.reshape(*batch_shape, N)


# Every launch goes to the current stream of the tensors' own device; make that device current for the duration of the
# call when the caller's current device is another one (multi-GPU processes).
for _name, _obj in list(globals().items()):
    if callable(_obj) and getattr(_obj, "__module__", None) == __name__ and not _name.startswith("_") \
            and not isinstance(_obj, type):
        globals()[_name] = device_guard(_obj)
AddedDiagPreconditioner.__init__ = device_guard(AddedDiagPreconditioner.__init__)
AddedDiagPreconditioner._apply = device_guard(AddedDiagPreconditioner._apply)

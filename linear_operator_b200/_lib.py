"""ctypes binding of ``liblob_b200.so`` (C ABI declared in ``include/lob_b200.h``).

PyTorch is only plumbing here: it owns device memory (``tensor.data_ptr()``) and the current CUDA stream.  Every
numerical step of the Krylov path is a call into the library below.  There is NO CPU fallback: asking for compute
on a non-CUDA tensor, or without the built library, raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_int32, c_int64, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LOB_LIB_PATH", os.path.join(_HERE, "csrc", "liblob_b200.so"))

F32, F64 = 0, 1
_DT = {torch.float32: F32, torch.float64: F64}


class LobError(RuntimeError):
    pass


class CgParams(Structure):
    _fields_ = [
        ("B", c_int64),
        ("N", c_int64),
        ("C", c_int64),
        ("dtype", c_int32),
        ("n_tridiag", c_int32),
        ("n_tridiag_iter", c_int32),
        ("max_iter", c_int32),
        ("n_iter", c_int32),
        ("has_precond", c_int32),
        ("tolerance", c_double),
        ("eps", c_double),
        ("stop_updating_after", c_double),
    ]


class CgStatus(Structure):
    _fields_ = [
        ("stop", c_int32),
        ("tolerance_reached", c_int32),
        ("iterations", c_int32),
        ("update_tridiag", c_int32),
        ("last_tridiag_iter", c_int32),
        ("nan_detected", c_int32),
        ("all_converged_at_start", c_int32),
        ("reserved", c_int32),
        ("residual_norm_mean", c_double),
    ]


_P = c_void_p
_SIGS = {
    "lob_version": (ctypes.c_int, []),
    "lob_last_error": (c_char_p, []),
    "lob_launch_count": (c_int64, []),
    "lob_cg_workspace_bytes": (c_size_t, [POINTER(CgParams)]),
    "lob_cg_setup": (ctypes.c_int, [POINTER(CgParams), _P, _P, _P, _P, _P, _P, _P]),
    "lob_cg_residual_init": (ctypes.c_int, [POINTER(CgParams), _P, _P, _P, _P, _P]),
    "lob_cg_direction_init": (ctypes.c_int, [POINTER(CgParams), _P, _P, _P, _P, _P, c_int32, _P]),
    "lob_cg_step_xr": (ctypes.c_int, [POINTER(CgParams), _P, c_int32, _P, _P, _P, _P, _P, c_int32, _P]),
    "lob_cg_step_p": (ctypes.c_int, [POINTER(CgParams), _P, c_int32, _P, _P, _P, _P, _P, c_int32, _P]),
    "lob_cg_poll_sync": (ctypes.c_int, [POINTER(CgParams), _P, POINTER(CgStatus), _P]),
    "lob_cg_finish": (ctypes.c_int, [POINTER(CgParams), _P, _P, _P]),
    "lob_debug_pin_dense_impl": (ctypes.c_int, [c_int32]),
    "lob_dense_matmul_parts": (c_int32, [c_int64]),
    "lob_dense_matmul_workspace_bytes": (c_size_t, [c_int32, c_int64, c_int64, c_int64, c_int64]),
    "lob_dense_matmul": (
        ctypes.c_int,
        [c_int32, c_int64, c_int64, c_int64, c_int64, _P, c_int64, c_int64, _P, _P, _P, c_int64, c_int64, _P, _P,
         c_size_t, _P],
    ),
    "lob_dense_matmul_ex": (
        ctypes.c_int,
        [c_int32, c_int64, c_int64, c_int64, c_int64, _P, c_int64, c_int64, _P, _P, _P, _P, c_int64, _P, c_int64,
         c_int64, _P, _P, c_size_t, _P],
    ),
    "lob_matmul_nn": (
        ctypes.c_int,
        [c_int32, c_int64, c_int64, c_int64, c_int64, _P, c_int64, c_int64, _P, c_int64, _P, c_double, _P],
    ),
    "lob_tn_matmul_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int64]),
    "lob_tn_matmul": (
        ctypes.c_int,
        [c_int32, c_int32, c_int64, c_int64, c_int64, c_int64, _P, c_int64, _P, c_int64, _P, _P, _P],
    ),
    "lob_lanczos_init": (ctypes.c_int, [c_int32, c_int64, c_int64, c_int64, _P, _P, _P]),
    "lob_lanczos_step": (
        ctypes.c_int,
        [c_int32, c_int32, c_int64, c_int64, c_int64, c_int32, c_int32, _P, _P, _P, _P, c_double, _P],
    ),
    "lob_pivchol_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int32]),
    "lob_pivchol_dense": (
        ctypes.c_int,
        [c_int32, c_int64, c_int64, c_int32, c_double, _P, c_int64, c_int64, _P, _P, _P, _P, _P],
    ),
    "lob_pivchol_kron": (
        ctypes.c_int,
        [c_int32, c_int64, c_int32, POINTER(c_int64), POINTER(c_void_p), POINTER(c_int64), c_int32, c_double, _P, _P,
         _P, _P, _P],
    ),
    "lob_pivchol_toeplitz": (
        ctypes.c_int,
        [c_int32, c_int64, c_int64, _P, c_int64, c_int32, c_double, _P, _P, _P, _P, _P],
    ),
    "lob_pivchol_rows_begin": (ctypes.c_int, [c_int32, c_int64, c_int64, c_int32, _P, _P, _P, _P]),
    "lob_pivchol_rows_pivot": (ctypes.c_int, [c_int32, c_int64, c_int64, c_int32, c_int32, c_double, _P, _P, _P, _P, _P]),
    "lob_pivchol_rows_update": (ctypes.c_int, [c_int32, c_int64, c_int64, c_int32, c_int32, _P, _P, _P, _P]),
    "lob_pivchol_rows_status": (ctypes.c_int, [c_int32, c_int64, c_int64, c_int32, _P, _P, _P, _P]),
    "lob_transpose_rows": (ctypes.c_int, [c_int32, c_int64, c_int64, c_int64, c_int64, _P, _P, _P]),
    "lob_precond_factor": (
        ctypes.c_int,
        [c_int32, c_int64, c_int32, _P, c_double, _P, c_int64, _P, _P, _P, _P, _P],
    ),
    "lob_precond_combine": (
        ctypes.c_int,
        [c_int32, c_int64, c_int64, c_int64, _P, _P, _P, c_int64, c_int64, c_int32, _P, _P],
    ),
    "lob_scale_rows": (ctypes.c_int, [c_int32, c_int64, c_int64, c_int64, _P, _P, c_int64, c_int64, c_int32, _P, _P]),
    "lob_colred_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "lob_probe_assemble": (
        ctypes.c_int,
        [c_int32, c_int64, c_int64, c_int64, _P, _P, _P, c_int64, c_int64, _P, _P, _P, _P],
    ),
    "lob_col_dots": (
        ctypes.c_int,
        [c_int32, c_int64, c_int64, c_int64, _P, c_int64, c_int64, _P, c_int64, c_int64, _P, _P, _P],
    ),
    "lob_tridiag_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int32, c_int32]),
    "lob_tridiag_eigh_slq": (
        ctypes.c_int,
        [c_int32, c_int64, c_int64, c_int32, c_int64, _P, _P, _P, _P, _P, _P, _P],
    ),
    "lob_kron_mode_matmul": (ctypes.c_int, [c_int32, c_int64, c_int64, c_int64, c_int64, _P, c_int64, _P, _P, _P]),
    "lob_toeplitz_pad": (ctypes.c_int, [c_int32, c_int64, c_int64, c_int64, c_int64, _P, _P, _P]),
    "lob_toeplitz_embed": (ctypes.c_int, [c_int32, c_int64, c_int64, c_int64, _P, c_int64, _P, _P]),
    "lob_toeplitz_mul": (ctypes.c_int, [c_int32, c_int64, c_int64, c_int64, _P, c_int64, _P, _P]),
    "lob_toeplitz_colmax": (ctypes.c_int, [c_int32, c_int64, c_int64, c_int64, _P, _P, _P]),
    "lob_toeplitz_pack": (ctypes.c_int, [c_int32, c_int64, c_int64, c_int64, c_int64, _P, _P, _P, _P]),
    "lob_toeplitz_mulr": (ctypes.c_int, [c_int32, c_int64, c_int64, c_int64, _P, c_int64, _P, _P]),
    "lob_toeplitz_unpack": (
        ctypes.c_int,
        [c_int32, c_int64, c_int64, c_int64, c_int64, _P, c_double, _P, _P, _P, c_int64, c_int64, _P, _P, _P],
    ),
    "lob_toeplitz_unpad": (
        ctypes.c_int,
        [c_int32, c_int64, c_int64, c_int64, c_int64, _P, c_double, _P, _P, c_int64, c_int64, _P, _P, _P],
    ),
    "lob_toeplitz_unpad_parts": (c_int32, [c_int32, c_int64, c_int64]),
    "lob_toeplitz_unpack_parts": (c_int32, [c_int32, c_int64, c_int64]),
    "lob_cap_solve_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int64]),
    "lob_cap_solve": (ctypes.c_int, [c_int32, c_int64, c_int32, c_int64, _P, c_int64, _P, _P, _P, _P, _P]),
    "lob_gemm3x_splits": (c_int32, [c_int64, c_int64, c_int64, c_int64, c_int32]),
    "lob_gemm3x_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int64, c_int32]),
    "lob_gemm3x": (
        ctypes.c_int,
        [c_int64, c_int64, c_int64, c_int64, _P, c_int32, c_int64, c_int64, c_int64, _P, c_int32, c_int64, c_int64,
         c_int64, _P, c_int32, c_int64, c_int64, c_double, _P, c_int64, _P, c_int64, c_int64, _P, c_int64, c_int32,
         c_int32, _P, c_size_t, _P],
    ),
    "lob_bilinear_dense": (ctypes.c_int, [c_int32, c_int64, c_int64, c_int64, c_int64, _P, _P, _P, _P, c_int32, _P]),
    "lob_bilinear_diag": (ctypes.c_int, [c_int32, c_int64, c_int64, c_int64, _P, _P, _P, _P, _P]),
    "lob_tri_inverse": (ctypes.c_int, [c_int32, c_int64, c_int32, _P, c_int64, c_int64, _P, _P]),
    "lob_toeplitz_cross_spectrum": (ctypes.c_int, [c_int32, c_int64, c_int64, c_int64, _P, _P, _P, _P, _P]),
    "lob_toeplitz_deriv_finish": (ctypes.c_int, [c_int32, c_int64, c_int64, c_int64, _P, c_double, _P, _P]),
    "lob_minres_z": (ctypes.c_int, [c_int32, c_int64, c_int64, c_int64, _P, _P, _P, _P, _P, _P]),
    "lob_minres_scalars": (
        ctypes.c_int,
        [c_int32, c_int64, c_int64, c_int64, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_double, _P],
    ),
    "lob_minres_update": (
        ctypes.c_int,
        [c_int32, c_int64, c_int64, c_int64, c_int64, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    ),
}

EXPORTED_SYMBOLS = tuple(_SIGS)

_lib = None


def load():
    """Loads the shared library (once).  Raises if it has not been built: the product has no other compute path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LobError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(linear_operator_b200 has no CPU or PyTorch fallback)."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


_DENSE_IMPLS = {None: 0, "auto": 0, "stream2": 1, "stream": 2, "tc": 3, "simt": 4, "stream2p": 5}


def pin_dense_impl(name=None):
    """Test hook: pins the fp32 dense-matmul kernel ("stream2p", "stream2", "stream", "tc", "simt"; None = automatic dispatch)."""
    check(load().lob_debug_pin_dense_impl(_DENSE_IMPLS[name]), "lob_debug_pin_dense_impl")


def launch_count() -> int:
    return int(load().lob_launch_count())


UNSUPPORTED = -3


def check(status: int, what: str):
    if status != 0:
        msg = load().lob_last_error()
        raise LobError(f"{what} failed ({status}): {msg.decode() if msg else ''}")


def dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise LobError(f"linear_operator_b200 supports float32 and float64 tensors, got {t.dtype}") from None


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise LobError(
                "linear_operator_b200 computes on CUDA tensors only (sm_100a kernels, no CPU fallback); "
                f"got a tensor on {t.device}."
            )


def _first_cuda_device(objs):
    for o in objs:
        if torch.is_tensor(o):
            if o.is_cuda:
                return o.device
        elif isinstance(o, (list, tuple)):
            d = _first_cuda_device(o)
            if d is not None:
                return d
        elif hasattr(o, "representation") and callable(o.representation):  # a LinearOperator
            d = _first_cuda_device(o.representation())
            if d is not None:
                return d
    return None


def device_guard(fn):
    """Runs ``fn`` with the device of its first CUDA tensor argument as the current device: the C ABI launches on the
    stream handle it is given, which belongs to that device."""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        dev = _first_cuda_device(args) or _first_cuda_device(tuple(kwargs.values()))
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)

    return wrapped


def ptr(t) -> c_void_p:
    return c_void_p(0 if t is None else t.data_ptr())


def stream(t: torch.Tensor) -> c_void_p:
    return c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)

"""Context-manager settings read by the Krylov hot path.

Same names, defaults and semantics as the reference's ``linear_operator/settings.py`` (process-global class state,
not thread-local; usable as ``with settings.cg_tolerance(1e-3):`` or queried with ``.value()`` / ``.on()``), restricted
to the knobs this path reads.  Reference lines: flags :58-93, values :96-118, dtype values :9-55.
"""
from __future__ import annotations

import logging

import torch


class _Flag:
    """Boolean switch: ``with flag(True):``, ``flag.on()``, ``flag.off()``  (reference settings.py:58-93)."""

    _default = False
    _state = None  # per-subclass

    def __init__(self, state: bool = True):
        self._enter_state = state
        self._prev = None

    @classmethod
    def is_default(cls):
        return cls._state is None or cls._state == cls._default

    @classmethod
    def on(cls) -> bool:
        return cls._default if cls._state is None else bool(cls._state)

    @classmethod
    def off(cls) -> bool:
        return not cls.on()

    @classmethod
    def _set_state(cls, state):
        cls._state = state

    def __enter__(self):
        self._prev = self.__class__._state
        self.__class__._set_state(self._enter_state)
        return self

    def __exit__(self, *exc):
        self.__class__._set_state(self._prev)
        return False


class _Value:
    """Scalar setting: ``with setting(v):``, ``setting.value()``  (reference settings.py:96-118)."""

    _global_value = None

    def __init__(self, value):
        self._enter_value = value
        self._prev = None

    @classmethod
    def value(cls):
        return cls._global_value

    @classmethod
    def _set_value(cls, value):
        cls._global_value = value

    def __enter__(self):
        self._prev = self.__class__.value()
        self.__class__._set_value(self._enter_value)
        return self

    def __exit__(self, *exc):
        self.__class__._set_value(self._prev)
        return False


class _DtypeValue:
    """Value with one slot per floating dtype: ``setting.value(dtype)``; ``with setting(float_value=..,
    double_value=.., half_value=..)``  (reference settings.py:9-55)."""

    _global_float_value = None
    _global_double_value = None
    _global_half_value = None

    def __init__(self, float_value=None, double_value=None, half_value=None):
        self._new = {"float": float_value, "double": double_value, "half": half_value}
        self._old = {}

    @classmethod
    def value(cls, dtype):
        if torch.is_tensor(dtype):
            dtype = dtype.dtype
        if dtype == torch.float:
            return cls._global_float_value
        if dtype == torch.double:
            return cls._global_double_value
        if dtype == torch.half:
            return cls._global_half_value
        raise RuntimeError(f"Unsupported dtype for {cls.__name__}.")

    @classmethod
    def _set_value(cls, float_value, double_value, half_value):
        if float_value is not None:
            cls._global_float_value = float_value
        if double_value is not None:
            cls._global_double_value = double_value
        if half_value is not None:
            cls._global_half_value = half_value

    def __enter__(self):
        cls = self.__class__
        self._old = {"float": cls._global_float_value, "double": cls._global_double_value,
                     "half": cls._global_half_value}
        cls._set_value(self._new["float"], self._new["double"], self._new["half"])
        return self

    def __exit__(self, *exc):
        cls = self.__class__
        cls._global_float_value = self._old["float"]
        cls._global_double_value = self._old["double"]
        cls._global_half_value = self._old["half"]
        return False


# ---- hot-path knobs (defaults: reference settings.py line cited) -------------------------------------------
class cg_tolerance(_Value):
    """Relative residual tolerance of CG; default 1 (:216-223)."""

    _global_value = 1


class minres_tolerance(_Value):
    """Relative update-term tolerance that terminates MINRES; default 1e-4 (reference settings.py:464-471)."""

    _global_value = 1e-4


class num_contour_quadrature(_Value):
    """Number of quadrature points of contour-integral quadrature; default 15 (reference settings.py:474-481)."""

    _global_value = 15


class max_cg_iterations(_Value):
    """Maximum CG iterations; default 1000 (:383-391)."""

    _global_value = 1000


class max_lanczos_quadrature_iterations(_Value):
    """Tridiagonal size for stochastic Lanczos quadrature; default 20 (:405-414)."""

    _global_value = 20


class max_cholesky_size(_Value):
    """Operators up to this size use dense Cholesky instead of CG; default 800 (:394-402)."""

    _global_value = 800


class max_preconditioner_size(_Value):
    """Rank of the pivoted-Cholesky preconditioner; default 15 (:417-425)."""

    _global_value = 15


class min_preconditioning_size(_Value):
    """Smallest operator that gets a preconditioner; default 2000 (:453-461)."""

    _global_value = 2000


class preconditioner_tolerance(_Value):
    """Early-stop tolerance of the pivoted Cholesky; default 1e-3 (:496-503)."""

    _global_value = 1e-3


class num_trace_samples(_Value):
    """Number of probe vectors for the log determinant; default 10 (:484-493)."""

    _global_value = 10


class max_root_decomposition_size(_Value):
    """Lanczos iterations for root decompositions; default 100."""

    _global_value = 100


class max_lanczos_iterations(_Value):
    _global_value = 100


class terminate_cg_by_size(_Flag):
    """Cap CG iterations at the operator size; default off (:534-541)."""

    _default = False


class skip_logdet_forward(_Flag):
    """Return zeros for the logdet in the forward pass; default off (:506-520)."""

    _default = False


class deterministic_probes(_Flag):
    """Deprecated in the reference (:245-262); kept so `with` blocks do not break.  Not implemented here."""

    _default = False
    probe_vectors = None


class ciq_samples(_Flag):
    """Contour-integral-quadrature sampling in ``zero_mean_mvn_samples`` (reference settings.py ciq_samples); default
    off (samples come from the Lanczos root decomposition)."""

    _default = False


class cuda_graphs(_Flag):
    """Not in the reference.  Small dense problems (the operator at most ``cuda_graphs.max_operator_bytes``) are
    launch-bound: ~90 kernel launches for a 21-iteration solve of N = 512.  With this flag on (default) linear_cg
    replays the whole fixed-length part of the solve as ONE CUDA graph on solver-owned static buffers (inputs are copied
    in, a few MB) and reads the control words once at the end; results are bit-identical to the eager launches."""

    _default = True
    max_operator_bytes = 32 << 20
    max_rhs_bytes = 8 << 20


class debug(_Flag):
    """Argument checking; default on (:265-275)."""

    _default = True


class memory_efficient(_Flag):
    _default = False


class trace_mode(_Flag):
    _default = False


class verbose_linalg(_Flag):
    """Logs one line per expensive routine (:587-605)."""

    _default = False
    logger = logging.getLogger("LinAlg (Verbose)")


class fast_computations:
    """``fast_computations(covar_root_decomposition=True, log_prob=True, solves=True)`` (:278-354)."""

    class covar_root_decomposition(_Flag):
        _default = True

    class log_prob(_Flag):
        _default = True

    class solves(_Flag):
        _default = True

    def __init__(self, covar_root_decomposition=True, log_prob=True, solves=True):
        self._ctx = [
            fast_computations.covar_root_decomposition(covar_root_decomposition),
            fast_computations.log_prob(log_prob),
            fast_computations.solves(solves),
        ]

    def __enter__(self):
        for c in self._ctx:
            c.__enter__()
        return self

    def __exit__(self, *exc):
        for c in reversed(self._ctx):
            c.__exit__(*exc)
        return False


class cholesky_jitter(_DtypeValue):
    """Jitter added when a Cholesky fails (:195-213)."""

    _global_float_value = 1e-6
    _global_double_value = 1e-8
    _global_half_value = 1e-3


class cholesky_max_tries(_Value):
    _global_value = 3


class tridiagonal_jitter(_Value):
    _global_value = 1e-6


class _linalg_dtype_symeig(_Value):
    _global_value = torch.double


class _linalg_dtype_cholesky(_Value):
    _global_value = torch.double


__all__ = [
    "cg_tolerance", "max_cg_iterations", "max_lanczos_quadrature_iterations", "max_cholesky_size",
    "max_preconditioner_size", "min_preconditioning_size", "preconditioner_tolerance", "num_trace_samples",
    "max_root_decomposition_size", "max_lanczos_iterations", "terminate_cg_by_size", "skip_logdet_forward",
    "deterministic_probes", "debug", "memory_efficient", "trace_mode", "verbose_linalg", "fast_computations",
    "cholesky_jitter", "cholesky_max_tries", "tridiagonal_jitter", "ciq_samples", "cuda_graphs", "minres_tolerance", "num_contour_quadrature",
]

"""``LinearOperator`` base class -- the drop-in boundary of the Krylov hot path.

Mirrors the part of the reference's ``operators/_linear_operator.py`` that the path needs (SURVEY.md section 8a row 19,
8b): construction / flattening (``_args``, ``representation``, ``representation_tree``, reference :149-163,:2076-2088),
the operator hooks subclasses override (``_matmul`` :170, ``_size`` :206, ``_transpose_nonbatch`` :213, ``_solve`` :781,
``_preconditioner`` :618, ``_solve_preconditioner`` :805, ``_probe_vectors_and_norms`` :629, ``_diagonal`` :563,
``_approx_diagonal`` :483, ``_get_indices`` :412, ``_expand_batch`` :395), the public methods that dispatch into the
path (``matmul`` :1844, ``solve`` :2324, ``inv_quad`` :1637, ``inv_quad_logdet`` :1688, ``logdet`` :1834,
``pivoted_cholesky`` :1975, ``add_diagonal`` :953, ``add_jitter`` :1001, ``zero_mean_mvn_samples`` :2746) and the
``__torch_function__`` routing by method *name* (:56-111,:2981-3009).

All arithmetic is delegated to ``liblob_b200`` through ``_kernels``; the composition algebra of the reference that the
north star does not name (Mul/Cat/Block/Interpolated/...) is out of scope and raises ``NotImplementedError``.
"""
from __future__ import annotations

import functools
import math
import numbers
import warnings
from collections import OrderedDict
from typing import Callable, Optional, Tuple

import torch
from torch import Tensor

from .. import _kernels, settings, utils
from ..utils.broadcasting import _matmul_broadcast_shape, _to_helper
from ..utils.memoize import cached
from ..utils.warnings import NumericalWarning
from .linear_operator_representation_tree import LinearOperatorRepresentationTree

_HANDLED_FUNCTIONS = {}
_HANDLED_SECOND_ARG_FUNCTIONS = {}


def _implements(torch_function: Callable) -> Callable:
    """Registers a method (by name, so subclass overrides win) as the handler of ``torch_function``."""

    def decorator(func):
        _HANDLED_FUNCTIONS[torch_function] = func.__name__
        return func

    return decorator


def _implements_second_arg(torch_function: Callable) -> Callable:
    """Handler for ``torch_function(tensor, linear_operator)``-style calls."""

    def decorator(func):
        _HANDLED_SECOND_ARG_FUNCTIONS[torch_function] = func.__name__
        return func

    return decorator


def _implements_symmetric(torch_function: Callable) -> Callable:
    def decorator(func):
        _HANDLED_FUNCTIONS[torch_function] = func.__name__
        _HANDLED_SECOND_ARG_FUNCTIONS[torch_function] = func.__name__
        return func

    return decorator


class LinearOperator(object):
    """A (batch of) matrices ``(*batch, M, N)`` defined by its action ``_matmul``.

    Subclass contract (reference ``docs/source/custom_linear_operators.rst:7-29``): implement ``_matmul(rhs)``,
    ``_size()`` and ``_transpose_nonbatch()``; every tensor / operator the instance is built from must be passed to
    ``super().__init__`` so it can be flattened and rebuilt by the autograd Functions.
    """

    def _check_args(self, *args, **kwargs) -> Optional[str]:
        return None

    def __init__(self, *args, **kwargs):
        if settings.debug.on():
            err = self._check_args(*args, **kwargs)
            if err is not None:
                raise ValueError(err)
        self._args = args
        self._differentiable_kwargs = OrderedDict()
        self._nondifferentiable_kwargs = dict()
        for name, val in sorted(kwargs.items()):
            if torch.is_tensor(val) or isinstance(val, LinearOperator):
                self._differentiable_kwargs[name] = val
            else:
                self._nondifferentiable_kwargs[name] = val

    # ------------------------------------------------------------------ required hooks
    def _matmul(self, rhs: Tensor) -> Tensor:
        raise NotImplementedError(f"The class {self.__class__.__name__} requires a _matmul function!")

    def _size(self) -> torch.Size:
        raise NotImplementedError(f"The class {self.__class__.__name__} requires a _size function!")

    def _transpose_nonbatch(self) -> "LinearOperator":
        raise NotImplementedError(f"The class {self.__class__.__name__} requires a _transpose_nonbatch function!")

    # ------------------------------------------------------------------ optional hooks with generic defaults
    @property
    def _kwargs(self):
        return {**self._differentiable_kwargs, **self._nondifferentiable_kwargs}

    def _diagonal(self) -> Tensor:
        """Diagonal through the operator's own rows (generic operators: one-hot products in row blocks)."""
        n = self.size(-1)
        out = torch.empty(*self.batch_shape, n, dtype=self.dtype, device=self.device)
        blk = 256
        for s in range(0, n, blk):
            e = min(s + blk, n)
            eye = torch.zeros(n, e - s, dtype=self.dtype, device=self.device)
            eye[s:e] = torch.eye(e - s, dtype=self.dtype, device=self.device)
            cols = self._matmul(eye.expand(*self.batch_shape, n, e - s))
            out[..., s:e] = cols[..., s:e, :].diagonal(dim1=-1, dim2=-2)
        return out

    def _approx_diagonal(self) -> Tensor:
        return self._diagonal()

    def _expand_batch(self, batch_shape) -> "LinearOperator":
        raise NotImplementedError(f"{self.__class__.__name__} does not implement _expand_batch")

    def _get_indices(self, row_index, col_index, *batch_indices) -> Tensor:
        raise NotImplementedError(f"{self.__class__.__name__} does not implement _get_indices")

    def _pivoted_cholesky(self, rank: int, error_tol: float):
        """(L (*batch, N, m), permutation (*batch, N) int64).  Default: rows through ``_get_indices`` (the reference's
        generic route, functions/_pivoted_cholesky.py:57-98); classes with a device row functor override this."""
        return _kernels.pivoted_cholesky_rows(self, rank, error_tol)

    def _bilinear_derivative(self, left_vecs: Tensor, right_vecs: Tensor) -> Tuple[Optional[Tensor], ...]:
        """d/d(theta) sum_i u_i^T K(theta) v_i for every tensor of ``representation()`` (reference :336-393).  The
        reference's default differentiates ``_matmul`` with autograd; the ``_matmul``s of this package are CUDA kernels
        outside autograd, so every operator class of the path carries its closed form and a user-defined operator that
        wants gradients overrides this hook (the reference's own extension point)."""
        if not any(a.requires_grad for a in self.representation()):
            return (None,) * len(self.representation())
        raise NotImplementedError(
            f"{self.__class__.__name__} must implement _bilinear_derivative(left_vecs, right_vecs) to be "
            "differentiated through the Krylov path."
        )

    def _preconditioner(self) -> Tuple[Optional[Callable], Optional["LinearOperator"], Optional[Tensor]]:
        """(closure M^-1 v, LinearOperator M, logdet M) or (None, None, None)  (reference :618-627)."""
        return None, None, None

    def _probe_vectors_and_norms(self):
        return None, None

    def _solve_preconditioner(self) -> Optional[Callable]:
        """Preconditioner used by plain solves: only the closure is needed (reference :805-848)."""
        base_precond, _, _ = self._preconditioner()
        return base_precond

    def _matmul_closure(self):
        """Closure handed to linear_cg; subclasses with a fused matmul attach ``closure.fused``."""
        return self._matmul

    def _solve(self, rhs: Tensor, preconditioner: Optional[Callable] = None, num_tridiag: int = 0):
        """mBCG on this operator (reference :781-803).  ``utils.linear_cg`` is looked up on the module at call time."""
        rhs = rhs.expand(*torch.broadcast_shapes(self.batch_shape, rhs.shape[:-2]), *rhs.shape[-2:])
        return utils.linear_cg(
            self._matmul_closure(),
            rhs,
            n_tridiag=num_tridiag,
            max_iter=settings.max_cg_iterations.value(),
            max_tridiag_iter=settings.max_lanczos_quadrature_iterations.value(),
            preconditioner=preconditioner,
            _skip_initial_matmul=True,
        )

    # ------------------------------------------------------------------ flattening for the Functions
    def representation(self) -> Tuple[Tensor, ...]:
        """Flat tuple of the leaf tensors that define this operator (reference :2076-2088)."""
        out = []
        for arg in list(self._args) + list(self._differentiable_kwargs.values()):
            if torch.is_tensor(arg):
                out.append(arg)
            elif hasattr(arg, "representation") and callable(arg.representation):
                out += list(arg.representation())
            else:
                raise RuntimeError(f"Representation of a LinearOperator should consist only of Tensors, got {type(arg)}")
        return tuple(out)

    def representation_tree(self) -> LinearOperatorRepresentationTree:
        return LinearOperatorRepresentationTree(self)

    # ------------------------------------------------------------------ shape / dtype / device
    @property
    def shape(self) -> torch.Size:
        return self._size()

    def size(self, dim: Optional[int] = None):
        s = self._size()
        return s if dim is None else s[dim]

    def dim(self) -> int:
        return len(self._size())

    def ndimension(self) -> int:
        return self.dim()

    def numel(self) -> int:
        return self.shape.numel()

    @property
    def batch_shape(self) -> torch.Size:
        return self.shape[:-2]

    @property
    def batch_dim(self) -> int:
        return len(self.batch_shape)

    @property
    def matrix_shape(self) -> torch.Size:
        return self.shape[-2:]

    @property
    def is_square(self) -> bool:
        return self.matrix_shape[0] == self.matrix_shape[1]

    @property
    def dtype(self) -> Optional[torch.dtype]:
        for a in self.representation():
            return a.dtype
        return self._nondifferentiable_kwargs.get("dtype", None)

    @property
    def device(self) -> Optional[torch.device]:
        for a in self.representation():
            return a.device
        return self._nondifferentiable_kwargs.get("device", None)

    @property
    def requires_grad(self) -> bool:
        return any(a.requires_grad for a in self.representation())

    def _rebuild(self, fn):
        args = [fn(a) if (torch.is_tensor(a) or isinstance(a, LinearOperator)) else a for a in self._args]
        kwargs = {k: fn(v) if (torch.is_tensor(v) or isinstance(v, LinearOperator)) else v
                  for k, v in self._kwargs.items()}
        return self.__class__(*args, **kwargs)

    def detach(self) -> "LinearOperator":
        return self._rebuild(lambda a: a.detach())

    def clone(self) -> "LinearOperator":
        return self._rebuild(lambda a: a.clone())

    def to(self, *args, **kwargs) -> "LinearOperator":
        device, dtype = _to_helper(*args, **kwargs)

        def conv(a):
            if isinstance(a, LinearOperator):
                return a.to(*args, **kwargs)
            return a.to(device=device, dtype=dtype if a.is_floating_point() else None)

        new_kwargs = {}
        for k, v in self._kwargs.items():
            if torch.is_tensor(v) or isinstance(v, LinearOperator):
                new_kwargs[k] = conv(v)
            elif k == "device" and device is not None:
                new_kwargs[k] = device
            elif k == "dtype" and dtype is not None:
                new_kwargs[k] = dtype
            else:
                new_kwargs[k] = v
        new_args = [conv(a) if (torch.is_tensor(a) or isinstance(a, LinearOperator)) else a for a in self._args]
        return self.__class__(*new_args, **new_kwargs)

    def cuda(self, device_id=None) -> "LinearOperator":
        return self.to(torch.device("cuda", device_id) if device_id is not None else torch.device("cuda"))

    def cpu(self) -> "LinearOperator":
        return self.to(torch.device("cpu"))

    def double(self) -> "LinearOperator":
        return self.to(torch.double)

    def float(self) -> "LinearOperator":
        return self.to(torch.float)

    def type(self, dtype) -> "LinearOperator":
        return self.to(dtype)

    def requires_grad_(self, val: bool) -> "LinearOperator":
        for a in self.representation():
            a.requires_grad_(val)
        return self

    # ------------------------------------------------------------------ algebra needed by the path
    @property
    def mT(self) -> "LinearOperator":
        return self._transpose_nonbatch()

    def t(self) -> "LinearOperator":
        if self.dim() != 2:
            raise RuntimeError("Cannot call t for more than 2 dimensions")
        return self._transpose_nonbatch()

    @_implements(torch.transpose)
    def transpose(self, dim1: int, dim2: int) -> "LinearOperator":
        nd = self.dim()
        dim1, dim2 = dim1 % nd, dim2 % nd
        if {dim1, dim2} == {nd - 2, nd - 1}:
            return self._transpose_nonbatch()
        raise NotImplementedError("Only the transpose of the two matrix dimensions is supported on this path.")

    def expand(self, *sizes) -> "LinearOperator":
        if len(sizes) == 1 and not isinstance(sizes[0], numbers.Integral):
            sizes = tuple(sizes[0])
        if tuple(sizes[-2:]) != tuple(self.matrix_shape) and tuple(sizes[-2:]) != (-1, -1):
            raise RuntimeError(f"Invalid expand arguments {sizes}: the matrix dimensions cannot change.")
        if torch.Size(sizes[:-2]) == self.batch_shape:
            return self
        return self._expand_batch(torch.Size(sizes[:-2]))

    @_implements(torch.matmul)
    def matmul(self, other):
        """``self @ other`` (reference :1844-1866 + functions/_matmul.py:8-66; forward only)."""
        if isinstance(other, LinearOperator):
            raise NotImplementedError("LinearOperator @ LinearOperator (MatmulLinearOperator) is outside the Krylov path.")
        _matmul_broadcast_shape(self.shape, other.shape)
        if other.ndimension() == 1:
            return self._matmul(other.unsqueeze(-1)).squeeze(-1)
        return self._matmul(other)

    def __matmul__(self, other):
        return self.matmul(other)

    @_implements_second_arg(torch.matmul)
    def rmatmul(self, other: Tensor) -> Tensor:
        if other.ndim == 1:
            return self.mT.matmul(other)
        return self.mT.matmul(other.mT).mT

    def __rmatmul__(self, other):
        return self.rmatmul(other)

    def to_dense(self) -> Tensor:
        """Explicit matrix, by multiplying with the identity in column blocks (reference :2521-2540)."""
        n = self.size(-1)
        eye = torch.eye(n, dtype=self.dtype, device=self.device).expand(*self.batch_shape, n, n)
        return self._matmul(eye.contiguous())

    @_implements(torch.diagonal)
    def diagonal(self, offset: int = 0, dim1: int = -2, dim2: int = -1) -> Tensor:
        if not (offset == 0 and ((dim1 == -2 and dim2 == -1) or (dim1 == -1 and dim2 == -2))):
            raise NotImplementedError("LinearOperator.diagonal only computes the diagonal of the last two dimensions.")
        if not self.is_square:
            raise RuntimeError("The diagonal is only defined for square operators.")
        return self._diagonal()

    def add_diagonal(self, diag: Tensor) -> "LinearOperator":
        """``self + diag_embed(diag)`` as an AddedDiagLinearOperator (reference :953-999)."""
        from .added_diag_linear_operator import AddedDiagLinearOperator
        from .diag_linear_operator import ConstantDiagLinearOperator, DiagLinearOperator

        if not self.is_square:
            raise RuntimeError("add_diagonal only defined for square matrices")
        diag_shape = diag.shape
        n = self.size(-1)
        if len(diag_shape) == 0:
            diag_op = ConstantDiagLinearOperator(diag.unsqueeze(-1), diag_shape=n)
        elif diag_shape[-1] == 1:
            diag_op = ConstantDiagLinearOperator(diag, diag_shape=n)
        else:
            try:
                expanded = diag.expand(*self.batch_shape, n) if len(diag_shape) <= len(self.shape) - 1 else diag
            except RuntimeError:
                raise RuntimeError(
                    "add_diagonal for LinearOperator of size {} received invalid diagonal of size {}.".format(
                        self.shape, diag_shape
                    )
                ) from None
            diag_op = DiagLinearOperator(expanded)
        return AddedDiagLinearOperator(self, diag_op)

    def add_jitter(self, jitter_val: float = 1e-3) -> "LinearOperator":
        """Adds ``jitter_val`` to the diagonal (reference :1001-1013)."""
        diag = torch.tensor(jitter_val, dtype=self.dtype, device=self.device)
        return self.add_diagonal(diag)

    @_implements_symmetric(torch.add)
    def __add__(self, other):
        from .added_diag_linear_operator import AddedDiagLinearOperator
        from .dense_linear_operator import to_linear_operator
        from .diag_linear_operator import DiagLinearOperator
        from .sum_linear_operator import SumLinearOperator

        if isinstance(other, numbers.Number) and other == 0:
            return self
        if isinstance(other, DiagLinearOperator):
            return AddedDiagLinearOperator(self, other)
        if isinstance(other, Tensor):
            other = to_linear_operator(other)
        if isinstance(other, LinearOperator):
            return SumLinearOperator(self, other)
        return NotImplemented

    def __radd__(self, other):
        return self + other

    def add(self, other, alpha=None):
        if alpha is not None and alpha != 1:
            raise NotImplementedError("add with alpha != 1 is outside the Krylov path.")
        return self + other

    # ------------------------------------------------------------------ the Krylov entry points
    def pivoted_cholesky(self, rank: int, error_tol: Optional[float] = None, return_pivots: bool = False):
        """Rank-``rank`` pivoted Cholesky factor ``L`` (*batch, N, m) (reference :1975-2003)."""
        from ..functions._pivoted_cholesky import PivotedCholesky

        res, pivots = PivotedCholesky.apply(self.representation_tree(), rank, error_tol, *self.representation())
        return (res, pivots) if return_pivots else res

    @cached(name="cholesky")
    def cholesky(self, upper: bool = False):
        """Dense Cholesky for operators below ``max_cholesky_size`` -- NOT the Krylov path (SURVEY.md section 3.1:
        "dense path, NOT hot"); provided so small problems dispatch like the reference (:1713-1731).  Uses
        torch.linalg (cuSOLVER) on the materialised matrix."""
        dense = self.to_dense()
        chol, info = torch.linalg.cholesky_ex(dense)
        if torch.any(info):
            from ..utils.errors import NotPSDError

            raise NotPSDError("Matrix not positive definite in the dense Cholesky path.")
        return chol.mT if upper else chol

    def _cholesky_inv_quad_logdet(self, inv_quad_rhs, logdet, reduce_inv_quad):
        chol = self.cholesky()
        inv_quad_term = None
        logdet_term = None
        if inv_quad_rhs is not None:
            rhs = inv_quad_rhs.unsqueeze(-1) if inv_quad_rhs.dim() == 1 else inv_quad_rhs
            half = torch.linalg.solve_triangular(chol, rhs, upper=False)
            inv_quad_term = half.pow(2).sum(-2)
            if reduce_inv_quad:
                inv_quad_term = inv_quad_term.sum(-1)
        if logdet:
            logdet_term = chol.diagonal(dim1=-1, dim2=-2).log().sum(-1).mul(2)
        return inv_quad_term, logdet_term

    @_implements(torch.linalg.solve)
    def solve(self, right_tensor: Tensor, left_tensor: Optional[Tensor] = None) -> Tensor:
        """``self^-1 right_tensor`` (reference :2324-2379 -> functions/_solve.py)."""
        from ..functions._solve import Solve

        if not self.is_square:
            raise RuntimeError(
                "solve only operates on (batches of) square (positive semi-definite) LinearOperators. "
                "Got a {} of size {}.".format(self.__class__.__name__, self.size())
            )
        if self.dim() == 2 and right_tensor.dim() == 1:
            if self.shape[-1] != right_tensor.numel():
                raise RuntimeError(
                    "LinearOperator (size={}) cannot be multiplied with right-hand-side Tensor (size={}).".format(
                        self.shape, right_tensor.shape
                    )
                )
        func = Solve
        if left_tensor is None:
            return func.apply(self.representation_tree(), False, right_tensor, *self.representation())
        return func.apply(self.representation_tree(), True, left_tensor, right_tensor, *self.representation())

    def inv_quad(self, inv_quad_rhs: Tensor, reduce_inv_quad: bool = True) -> Tensor:
        """tr(R^T A^-1 R) (or its diagonal)  (reference :1637-1686 -> functions/_inv_quad.py)."""
        from ..functions._inv_quad import InvQuad

        if not self.is_square:
            raise RuntimeError(
                "inv_quad only operates on (batches of) square (positive semi-definite) LinearOperators. "
                "Got a {} of size {}.".format(self.__class__.__name__, self.size())
            )
        try:
            result_shape = _matmul_broadcast_shape(self.shape, inv_quad_rhs.shape)
        except RuntimeError:
            raise RuntimeError(
                "LinearOperator (size={}) cannot be multiplied with right-hand-side Tensor (size={}).".format(
                    self.shape, inv_quad_rhs.shape
                )
            ) from None
        args = (inv_quad_rhs.expand(*result_shape[:-2], *inv_quad_rhs.shape[-2:]),) + self.representation()
        inv_quad_term = InvQuad.apply(self.representation_tree(), *args)
        if reduce_inv_quad:
            inv_quad_term = inv_quad_term.sum(-1)
        return inv_quad_term

    def inv_quad_logdet(self, inv_quad_rhs: Optional[Tensor] = None, logdet: bool = False,
                        reduce_inv_quad: bool = True):
        """Inverse quadratic form and log determinant in one preconditioned mBCG run (reference :1688-1804)."""
        from ..functions._inv_quad_logdet import InvQuadLogdet
        from .identity_linear_operator import IdentityLinearOperator

        # small problems: dense Cholesky, like the reference (:1713-1731)
        if settings.fast_computations.log_prob.off() or (self.size(-1) <= settings.max_cholesky_size.value()):
            return self._cholesky_inv_quad_logdet(inv_quad_rhs, logdet, reduce_inv_quad)

        if not logdet:  # :1734-1739
            if inv_quad_rhs is None:
                raise RuntimeError("Either `inv_quad_rhs` or `logdet` must be specifed.")
            return self.inv_quad(inv_quad_rhs, reduce_inv_quad=reduce_inv_quad), torch.zeros(
                [], dtype=self.dtype, device=self.device
            )

        if not self.is_square:
            raise RuntimeError(
                "inv_quad_logdet only operates on (batches of) square (positive semi-definite) LinearOperators. "
                "Got a {} of size {}.".format(self.__class__.__name__, self.size())
            )
        if inv_quad_rhs is not None:  # :1749-1767
            if self.dim() == 2 and inv_quad_rhs.dim() == 1:
                if self.shape[-1] != inv_quad_rhs.numel():
                    raise RuntimeError(
                        "LinearOperator (size={}) cannot be multiplied with right-hand-side Tensor (size={}).".format(
                            self.shape, inv_quad_rhs.shape
                        )
                    )
            elif self.dim() != inv_quad_rhs.dim():
                raise RuntimeError(
                    "LinearOperator (size={}) and right-hand-side Tensor (size={}) should have the same number "
                    "of dimensions.".format(self.shape, inv_quad_rhs.shape)
                )
            elif self.batch_shape != inv_quad_rhs.shape[:-2] or self.shape[-1] != inv_quad_rhs.shape[-2]:
                raise RuntimeError(
                    "LinearOperator (size={}) cannot be multiplied with right-hand-side Tensor (size={}).".format(
                        self.shape, inv_quad_rhs.shape
                    )
                )

        args = self.representation()
        if inv_quad_rhs is not None:
            args = [inv_quad_rhs] + list(args)

        preconditioner, precond_lt, logdet_p = self._preconditioner()  # :1773
        if precond_lt is None:
            precond_lt = IdentityLinearOperator(
                diag_shape=self.size(-1), batch_shape=self.batch_shape, dtype=self.dtype, device=self.device
            )
            logdet_p = 0.0
        precond_args = precond_lt.representation()
        probe_vectors, probe_vector_norms = self._probe_vectors_and_norms()

        inv_quad_term, pinvk_logdet = InvQuadLogdet.apply(
            self.representation_tree(),
            precond_lt.representation_tree(),
            preconditioner,
            len(precond_args),
            (inv_quad_rhs is not None),
            probe_vectors,
            probe_vector_norms,
            *(list(args) + list(precond_args)),
        )
        logdet_term = pinvk_logdet + logdet_p  # :1799-1800
        if inv_quad_term.numel() and reduce_inv_quad:
            inv_quad_term = inv_quad_term.sum(-1)
        return inv_quad_term, logdet_term

    # ------------------------------------------------------------------ Lanczos decompositions
    def _root_decomposition_size(self) -> int:  # :715-721
        return settings.max_root_decomposition_size.value()

    def _choose_root_method(self) -> str:  # :543-561 (no decomposition caches on this path)
        if self.size(-1) <= settings.max_cholesky_size.value() or settings.fast_computations.covar_root_decomposition.off():
            return "cholesky"
        return "lanczos"

    def _root_decomposition(self):  # :689-713
        from ..functions._root_decomposition import RootDecomposition

        res, _ = RootDecomposition.apply(
            self.representation_tree(), self._root_decomposition_size(), self.dtype, self.device, self.batch_shape,
            self.matrix_shape, True, False, None, *self.representation(),
        )
        return res

    def _root_inv_decomposition(self, initial_vectors=None, test_vectors=None):  # :723-761
        from ..functions._root_decomposition import RootDecomposition

        _, inv_roots = RootDecomposition.apply(
            self.representation_tree(), self._root_decomposition_size(), self.dtype, self.device, self.batch_shape,
            self.matrix_shape, True, True, initial_vectors, *self.representation(),
        )
        return inv_roots

    def root_decomposition(self, method: Optional[str] = None) -> "LinearOperator":
        """RootLinearOperator R with R R^T ~ A (reference :2158-2218).  Methods on this path: "lanczos" (default above
        ``max_cholesky_size``), "cholesky", "pivoted_cholesky", "diagonalization"."""
        from .root_linear_operator import RootLinearOperator

        if not self.is_square:
            raise RuntimeError(
                "root_decomposition only operates on (batches of) square (symmetric) LinearOperators. "
                "Got a {} of size {}.".format(self.__class__.__name__, self.size())
            )
        if self.shape[-2:].numel() == 1:
            return RootLinearOperator(self.to_dense().sqrt())
        if method is None:
            method = self._choose_root_method()
        if method == "cholesky":
            return RootLinearOperator(self.cholesky())
        if method == "pivoted_cholesky":
            return RootLinearOperator(self.pivoted_cholesky(rank=self._root_decomposition_size()))
        if method == "diagonalization":
            evals, evecs = self.diagonalization()
            return RootLinearOperator(evecs.to_dense() * evals.clamp_min(0.0).sqrt().unsqueeze(-2))
        if method == "lanczos":
            return RootLinearOperator(self._root_decomposition())
        raise RuntimeError(f"Unknown root decomposition method '{method}'")

    def root_inv_decomposition(self, initial_vectors=None, test_vectors=None, method: Optional[str] = None):
        """RootLinearOperator R with R R^T ~ A^-1 (reference :2221-2310); "lanczos" with a single initial vector,
        "cholesky" and "diagonalization" are on this path."""
        from .root_linear_operator import RootLinearOperator

        if not self.is_square:
            raise RuntimeError(
                "root_inv_decomposition only operates on (batches of) square (symmetric) LinearOperators. "
                "Got a {} of size {}.".format(self.__class__.__name__, self.size())
            )
        if self.shape[-2:].numel() == 1:
            return RootLinearOperator(1 / self.to_dense().sqrt())
        if method is None:
            method = self._choose_root_method()
        if method == "cholesky":
            L = self.cholesky()
            eye = torch.eye(L.shape[-2], device=L.device, dtype=L.dtype)
            return RootLinearOperator(torch.linalg.solve_triangular(L, eye, upper=False).mT)
        if method == "lanczos":
            if initial_vectors is not None:
                if self.dim() == 2 and initial_vectors.dim() == 1:
                    if self.shape[-1] != initial_vectors.numel():
                        raise RuntimeError(
                            "LinearOperator (size={}) cannot be multiplied with initial_vectors (size={}).".format(
                                self.shape, initial_vectors.shape
                            )
                        )
                elif self.dim() != initial_vectors.dim():
                    raise RuntimeError(
                        "LinearOperator (size={}) and initial_vectors (size={}) should have the same number "
                        "of dimensions.".format(self.shape, initial_vectors.shape)
                    )
                elif self.batch_shape != initial_vectors.shape[:-2] or self.shape[-1] != initial_vectors.shape[-2]:
                    raise RuntimeError(
                        "LinearOperator (size={}) cannot be multiplied with initial_vectors (size={}).".format(
                            self.shape, initial_vectors.shape
                        )
                    )
                if initial_vectors.size(-1) > 1:
                    raise NotImplementedError(
                        "root_inv_decomposition with several initial vectors (test-vector selection, reference "
                        ":2311-2350) is outside the Krylov path."
                    )
            return RootLinearOperator(self._root_inv_decomposition(initial_vectors))
        if method == "diagonalization":
            evals, evecs = self.diagonalization()
            return RootLinearOperator(evecs.to_dense() * evals.clamp_min(1e-7).reciprocal().sqrt().unsqueeze(-2))
        raise RuntimeError(f"Unknown root inv decomposition method '{method}'")

    def diagonalization(self, method: Optional[str] = None):
        """(eigenvalues (*b, k), eigenvectors as an operator (*b, N, k)) with Q S Q^T ~ A (reference :1439-1482)."""
        from ..functions._diagonalization import Diagonalization
        from .dense_linear_operator import to_linear_operator

        if not self.is_square:
            raise RuntimeError(
                "diagonalization only operates on (batches of) square (symmetric) LinearOperators. "
                "Got a {} of size {}.".format(self.__class__.__name__, self.size())
            )
        if method is None:
            method = "symeig" if self.size(-1) <= settings.max_cholesky_size.value() else "lanczos"
        if method == "lanczos":
            evals, evecs = Diagonalization.apply(
                self.representation_tree(), self.device, self.dtype, self.matrix_shape,
                self._root_decomposition_size(), self.batch_shape, *self.representation(),
            )
            return evals, to_linear_operator(evecs)
        if method == "symeig":  # small dense problems, off the Krylov path like the dense Cholesky branch
            evals, evecs = torch.linalg.eigh(self.to_dense())
            return evals, to_linear_operator(evecs)
        raise RuntimeError(f"Unknown diagonalization method '{method}'")

    @_implements(torch.logdet)
    def logdet(self) -> Tensor:
        """log |A| (reference :1834-1842)."""
        _, res = self.inv_quad_logdet(inv_quad_rhs=None, logdet=True)
        return res

    def zero_mean_mvn_samples(self, num_samples: int) -> Tensor:
        """Samples from N(0, self): (num_samples, *batch, N) = R eps with R the root decomposition (Lanczos above
        ``max_cholesky_size``) and eps = randn(*batch, k, S)  (reference :2746-2793, non-CIQ branch)."""
        if settings.ciq_samples.on():  # :2758-2777: K^{1/2} eps through contour-integral quadrature + shifted MINRES
            base_samples = torch.randn(*self.batch_shape, self.size(-1), num_samples, dtype=self.dtype,
                                       device=self.device)
            base_samples = base_samples.permute(-1, *range(self.dim() - 1)).contiguous().unsqueeze(-1)
            solves, weights, _, _ = utils.contour_integral_quad(
                self, base_samples, inverse=False, num_contour_quadrature=settings.num_contour_quadrature.value())
            return (solves * weights).sum(0).squeeze(-1)
        if self.size()[-2:] == torch.Size([1, 1]):
            covar_root = self.to_dense().sqrt()
        else:
            covar_root = self.root_decomposition().root.to_dense()
        base_samples = torch.randn(*self.batch_shape, covar_root.size(-1), num_samples, dtype=self.dtype,
                                   device=self.device)
        if covar_root.is_cuda:
            samples = _kernels.matmul_nn(covar_root, base_samples)
        else:
            samples = covar_root.matmul(base_samples)
        return samples.permute(-1, *range(self.dim() - 1)).contiguous()

    # ------------------------------------------------------------------ indexing (the subset the path needs)
    def __getitem__(self, index):
        """Tensor-index gathers ``op[(*batch_idx, row_idx, col_idx)]`` as used by apply_permutation
        (utils/permutation.py:76-87) are routed to ``_get_indices``; integer / slice *batch* indexing rebuilds the
        operator from indexed leaves where the subclass supports it."""
        if not isinstance(index, tuple):
            index = (index,)
        if len(index) == self.dim() and all(torch.is_tensor(i) for i in index):
            *batch_idx, row_idx, col_idx = index
            shape = torch.broadcast_shapes(*[i.shape for i in index])
            batch_idx = [i.expand(shape) for i in batch_idx]
            return self._get_indices(row_idx.expand(shape), col_idx.expand(shape), *batch_idx)
        return self._getitem(index)

    def _getitem(self, index):
        raise NotImplementedError(f"{self.__class__.__name__} does not support this kind of indexing on the Krylov path.")

    # ------------------------------------------------------------------ torch-namespace routing
    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        """Routes ``torch.f(op, ...)`` to the method registered for ``f`` -- by *name*, so overrides in subclasses win
        (reference :2981-3009)."""
        if kwargs is None:
            kwargs = {}
        if not isinstance(args[0], cls):
            if func not in _HANDLED_SECOND_ARG_FUNCTIONS or not all(
                issubclass(t, (torch.Tensor, LinearOperator)) for t in types
            ):
                name = func.__name__.replace("linalg_", "linalg.")
                arg_classes = ", ".join(arg.__class__.__name__ for arg in args)
                kwarg_classes = ", ".join(f"{key}={val.__class__.__name__}" for key, val in kwargs.items())
                raise NotImplementedError(f"torch.{name}({arg_classes}, {kwarg_classes}) is not implemented.")
            # second-arg functions take the operator second: call handler(op, tensor)
            name = _HANDLED_SECOND_ARG_FUNCTIONS[func]
            return getattr(args[1].__class__, name)(args[1], args[0], *args[2:], **kwargs)
        if func not in _HANDLED_FUNCTIONS or not all(issubclass(t, (torch.Tensor, LinearOperator)) for t in types):
            name = func.__name__.replace("linalg_", "linalg.")
            arg_classes = ", ".join(arg.__class__.__name__ for arg in args)
            kwarg_classes = ", ".join(f"{key}={val.__class__.__name__}" for key, val in kwargs.items())
            raise NotImplementedError(f"torch.{name}({arg_classes}, {kwarg_classes}) is not implemented.")
        name = _HANDLED_FUNCTIONS[func]
        return getattr(args[0].__class__, name)(*args, **kwargs)

    def __repr__(self):
        return f"<{self.__class__.__module__}.{self.__class__.__name__} object of size {tuple(self.shape)}>"


def _solve_second_arg(op, rhs):  # torch.linalg.solve(op, rhs) routes through LinearOperator.solve
    return op.solve(rhs)


def to_dense(obj) -> Tensor:
    """Tensor -> itself, LinearOperator -> ``to_dense()`` (reference :3023-3034)."""
    if torch.is_tensor(obj):
        return obj
    if isinstance(obj, LinearOperator):
        return obj.to_dense()
    raise TypeError("object of class {} cannot be made into a Tensor".format(obj.__class__.__name__))


__all__ = ["LinearOperator", "to_dense"]

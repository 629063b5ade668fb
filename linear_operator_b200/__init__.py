"""linear_operator_b200 -- B200-native batched Krylov hot path behind the ``linear_operator`` API.

Drop-in for the path ``LinearOperator.inv_quad_logdet / solve / inv_quad / logdet / pivoted_cholesky`` of
cornellius-gp/linear_operator for PSD operators built from Dense / Diag / AddedDiag / Kronecker / Toeplitz /
(LowRank)Root operators: ``import linear_operator_b200 as linear_operator``.  All arithmetic runs in hand-written
sm_100a kernels (``csrc/``) reached through a C ABI (``include/lob_b200.h``); there is no CPU fallback.
"""
from . import operators, settings, utils
from .functions import (add_diagonal, add_jitter, diagonal, inv_quad, inv_quad_logdet, logdet, matmul,
                        pivoted_cholesky, solve)
from .operators import LinearOperator, to_dense, to_linear_operator

__version__ = "0.1.0"

__all__ = [
    "LinearOperator", "add_diagonal", "add_jitter", "diagonal", "inv_quad", "inv_quad_logdet", "logdet", "matmul",
    "operators", "pivoted_cholesky", "settings", "solve", "to_dense", "to_linear_operator", "utils",
]

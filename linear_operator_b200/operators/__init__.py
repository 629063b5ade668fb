"""Operator classes of the Krylov hot path.  Names and import paths mirror
``linear_operator.operators``."""
from ._linear_operator import LinearOperator, to_dense
from .added_diag_linear_operator import AddedDiagLinearOperator
from .dense_linear_operator import DenseLinearOperator, to_linear_operator
from .diag_linear_operator import ConstantDiagLinearOperator, DiagLinearOperator
from .identity_linear_operator import IdentityLinearOperator
from .kronecker_product_linear_operator import KroneckerProductLinearOperator
from .kronecker_product_added_diag_linear_operator import KroneckerProductAddedDiagLinearOperator
from .linear_operator_representation_tree import LinearOperatorRepresentationTree
from .low_rank_root_added_diag_linear_operator import LowRankRootAddedDiagLinearOperator
from .root_linear_operator import LowRankRootLinearOperator, RootLinearOperator
from .sum_linear_operator import PsdSumLinearOperator, SumLinearOperator
from .toeplitz_linear_operator import ToeplitzLinearOperator

__all__ = [
    "to_dense",
    "to_linear_operator",
    "AddedDiagLinearOperator",
    "ConstantDiagLinearOperator",
    "DenseLinearOperator",
    "DiagLinearOperator",
    "IdentityLinearOperator",
    "KroneckerProductAddedDiagLinearOperator",
    "KroneckerProductLinearOperator",
    "LinearOperator",
    "LinearOperatorRepresentationTree",
    "LowRankRootAddedDiagLinearOperator",
    "LowRankRootLinearOperator",
    "PsdSumLinearOperator",
    "RootLinearOperator",
    "SumLinearOperator",
    "ToeplitzLinearOperator",
]

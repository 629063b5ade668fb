"""Krylov utilities.  ``linear_cg`` is a module attribute looked up at call time by ``LinearOperator._solve`` so that
``unittest.mock.patch("linear_operator_b200.utils.linear_cg")`` works like it does for the reference
(operators/_linear_operator.py:796; linear_operator/test/linear_operator_test_case.py:555-556)."""
from . import broadcasting, errors, lanczos, memoize, stochastic_lq, warnings
from .contour_integral_quad import contour_integral_quad
from .linear_cg import linear_cg
from .minres import minres
from .stochastic_lq import StochasticLQ

__all__ = ["broadcasting", "contour_integral_quad", "errors", "lanczos", "linear_cg", "minres", "memoize", "stochastic_lq", "StochasticLQ", "warnings"]

"""Functional API of the hot path (reference: functions/__init__.py:17-271): thin wrappers that turn the input into a
LinearOperator and call the method."""
from __future__ import annotations

import torch


def _op(input):
    from ..operators import to_linear_operator

    return to_linear_operator(input)


def add_diagonal(input, diag):
    return _op(input).add_diagonal(diag)


def add_jitter(input, jitter_val: float = 1e-3):
    if hasattr(input, "add_jitter"):
        return input.add_jitter(jitter_val)
    return _op(input).add_jitter(jitter_val)


def inv_quad(input, inv_quad_rhs, reduce_inv_quad: bool = True):
    return _op(input).inv_quad(inv_quad_rhs, reduce_inv_quad=reduce_inv_quad)


def inv_quad_logdet(input, inv_quad_rhs=None, logdet: bool = False, reduce_inv_quad: bool = True):
    return _op(input).inv_quad_logdet(inv_quad_rhs=inv_quad_rhs, logdet=logdet, reduce_inv_quad=reduce_inv_quad)


def pivoted_cholesky(input, rank: int, error_tol=None, return_pivots: bool = False):
    return _op(input).pivoted_cholesky(rank=rank, error_tol=error_tol, return_pivots=return_pivots)


def solve(input, rhs, lhs=None):
    return _op(input).solve(right_tensor=rhs, left_tensor=lhs)


def matmul(input, other):
    return _op(input).matmul(other)


def logdet(input):
    return _op(input).logdet()


def diagonal(input):
    return _op(input).diagonal()


__all__ = ["add_diagonal", "add_jitter", "diagonal", "inv_quad", "inv_quad_logdet", "logdet", "matmul",
           "pivoted_cholesky", "solve"]

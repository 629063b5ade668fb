"""Shape helpers (reference: utils/broadcasting.py:7-42)."""
from __future__ import annotations

import torch


def _matmul_broadcast_shape(shape_a, shape_b, error_msg=None) -> torch.Size:
    """Shape of ``A @ B`` for ``A`` of ``shape_a`` and a matrix (or vector) ``B`` of ``shape_b``."""
    shape_a, shape_b = tuple(shape_a), tuple(shape_b)
    m, n = shape_a[-2:]
    if len(shape_b) == 1:
        if n != shape_b[-1]:
            raise RuntimeError(error_msg or f"Incompatible dimensions for matmul: {shape_a} and {shape_b}")
        return torch.Size(shape_a[:-1])
    if n != shape_b[-2]:
        raise RuntimeError(error_msg or f"Incompatible dimensions for matmul: {shape_a} and {shape_b}")
    try:
        batch = torch.broadcast_shapes(shape_a[:-2], shape_b[:-2])
    except RuntimeError:
        raise RuntimeError(error_msg or f"Batch shapes {shape_a[:-2]} and {shape_b[:-2]} do not broadcast") from None
    return torch.Size(tuple(batch) + (m, shape_b[-1]))


def _to_helper(*args, **kwargs):
    """Parses ``.to()`` arguments into (device, dtype) (reference: utils/generic.py)."""
    device, dtype = None, None
    for a in list(args) + list(kwargs.values()):
        if isinstance(a, torch.dtype):
            dtype = a
        elif isinstance(a, (torch.device, str, int)):
            device = torch.device(a)
        elif torch.is_tensor(a):
            device, dtype = a.device, a.dtype
    return device, dtype

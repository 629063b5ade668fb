#!/usr/bin/env python3
"""Reference fixtures for SURVEY 8f rank 2: RootDecomposition (root and inverse root through Lanczos, forward and
backward) with SUPPLIED initial vectors (the reference draws them with torch.randn otherwise, and the CPU and CUDA RNG
streams differ).  cd /tmp && python /root/repo/tests/golden/make_golden_round2_lanczos.py"""
import os
import sys
import warnings

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from linear_operator import settings  # noqa: E402
from linear_operator.functions._root_decomposition import RootDecomposition  # noqa: E402
from linear_operator.operators import AddedDiagLinearOperator, DenseLinearOperator, DiagLinearOperator  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
warnings.simplefilter("ignore")


def npy(t):
    return t.detach().cpu().numpy()


g = torch.Generator().manual_seed(81)
dt = torch.float64
n, k = 60, 12
w = torch.randn(2, n, 25, dtype=dt, generator=g)
a = (w @ w.mT / 25).requires_grad_(True)
d = (0.3 + torch.rand(2, n, dtype=dt, generator=g)).requires_grad_(True)
init = torch.randn(2, n, 1, dtype=dt, generator=g)
w_root = torch.randn(2, n, k, dtype=dt, generator=g)
w_inv = torch.randn(2, n, k, dtype=dt, generator=g)
op = AddedDiagLinearOperator(DenseLinearOperator(a), DiagLinearOperator(d))
root, inv = RootDecomposition.apply(op.representation_tree(), k, op.dtype, op.device, op.batch_shape, op.matrix_shape,
                                    True, True, init, *op.representation())
((root * w_root).sum() + (inv * w_inv).sum()).backward()
np.savez_compressed(os.path.join(OUT, "root_decomposition_f64.npz"), A=npy(a), d=npy(d), init=npy(init), max_iter=k,
                    root=npy(root), inv_root=npy(inv), w_root=npy(w_root), w_inv=npy(w_inv), grad_A=npy(a.grad),
                    grad_d=npy(d.grad))
print("wrote root_decomposition_f64", root.shape, inv.shape)

# ---- Kronecker + constant diagonal through the eigen path (kronecker_product_added_diag_...:51-224)
from linear_operator.operators import KroneckerProductLinearOperator  # noqa: E402

for tag, kdt in (("f64", torch.float64), ("f32", torch.float32)):
    fs = []
    for m in (5, 6, 7):
        wf = torch.randn(2, m, m, dtype=kdt, generator=g)
        fs.append(wf @ wf.mT / m + 0.1 * torch.eye(m, dtype=kdt))
    op = KroneckerProductLinearOperator(*fs).add_jitter(0.5)
    assert type(op).__name__ == "KroneckerProductAddedDiagLinearOperator"
    rhs = torch.randn(2, 210, 3, dtype=kdt, generator=g)
    with settings.max_cholesky_size(0):
        sol = op.solve(rhs)
        iq, ld = op.inv_quad_logdet(rhs, logdet=True)
        ld_only = op.logdet()
    np.savez_compressed(os.path.join(OUT, f"kron_added_diag_{tag}.npz"), f0=npy(fs[0]), f1=npy(fs[1]), f2=npy(fs[2]),
                        rhs=npy(rhs), jitter=0.5, solve=npy(sol), inv_quad=npy(iq), logdet=npy(ld), logdet_only=npy(ld_only))
    print("wrote kron_added_diag_" + tag, sol.shape, iq.shape, ld.shape)

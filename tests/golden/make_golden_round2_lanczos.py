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

# ---- shifted MINRES and contour-integral quadrature (utils/minres.py, utils/contour_integral_quad.py)
from linear_operator.utils.contour_integral_quad import contour_integral_quad  # noqa: E402
from linear_operator.utils.minres import minres  # noqa: E402

wm = torch.randn(2, 50, 50, dtype=dt, generator=g)
am = wm @ wm.mT / 50 + 0.5 * torch.eye(50, dtype=dt)
bm = torch.randn(2, 50, 3, dtype=dt, generator=g)
bm[..., 1] = 0.0  # a zero right-hand-side column (masked to zero in the output, minres.py:197)
sh = torch.tensor([0.0, 0.3, 2.0], dtype=dt)
minv = 1.0 / am.diagonal(dim1=-1, dim2=-2)
sol = minres(am.matmul, bm, shifts=sh, max_iter=100)
sol_pre = minres(am.matmul, bm, shifts=sh, max_iter=100, preconditioner=lambda v: v * minv.unsqueeze(-1))
sol_neg = minres(am.matmul, bm, shifts=-sh - 0.1, value=-1, max_iter=100)
sol_vec = minres(am[0].matmul, bm[0, :, 0], max_iter=100)
np.savez_compressed(os.path.join(OUT, "minres_f64.npz"), A=npy(am), rhs=npy(bm), shifts=npy(sh), minv=npy(minv),
                    solve=npy(sol), solve_precond=npy(sol_pre), solve_neg=npy(sol_neg), solve_vec=npy(sol_vec))
print("wrote minres_f64", sol.shape, sol_vec.shape)
am32, bm32 = am.float(), bm.float()
sol32 = minres(am32.matmul, bm32, shifts=sh.float(), max_iter=100)
np.savez_compressed(os.path.join(OUT, "minres_f32.npz"), A=npy(am32), rhs=npy(bm32), shifts=npy(sh.float()), solve=npy(sol32))
opq = AddedDiagLinearOperator(DenseLinearOperator(am), DiagLinearOperator(torch.full((2, 50), 0.2, dtype=dt)))
rq = torch.randn(2, 50, 2, dtype=dt, generator=g)
out = {}
for inv in (True, False):
    solves, weights, no_shift, shifts = contour_integral_quad(opq, rq, inverse=inv)
    tag = "inv" if inv else "sqrt"
    out.update({f"solves_{tag}": npy(solves), f"weights_{tag}": npy(weights), f"no_shift_{tag}": npy(no_shift),
                f"shifts_{tag}": npy(shifts), f"result_{tag}": npy((solves * weights).sum(0))})
np.savez_compressed(os.path.join(OUT, "ciq_f64.npz"), A=npy(am), d=np.full((2, 50), 0.2), rhs=npy(rq), **out)
print("wrote ciq_f64", {k: v.shape for k, v in out.items()})

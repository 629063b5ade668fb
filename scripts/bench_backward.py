"""Backward pass at BASELINE config 2's operator shape (N = 5000, 33 columns, rank-100 preconditioner) on a batch that
leaves room for the (B, N, N) gradient: the rank-C outer-product kernel alone, and the whole forward + backward call."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linear_operator_b200 import _kernels, settings  # noqa: E402
from linear_operator_b200.operators import AddedDiagLinearOperator, DenseLinearOperator, DiagLinearOperator  # noqa: E402

dev = torch.device("cuda:0")
B, N, S = int(os.environ.get("BATCH", 384)), 5000, 32
gen = torch.Generator(device=dev).manual_seed(1234)
sc = torch.logspace(0, -1.5, 256, device=dev)
sc = sc / sc.norm()
K = torch.empty(B, N, N, device=dev)
for s0 in range(0, B, 32):
    W = torch.randn(min(32, B - s0), N, 256, device=dev, generator=gen) * sc
    torch.bmm(W, W.mT, out=K[s0:s0 + W.shape[0]])
del W
d = torch.full((B, N), 0.5, device=dev)
rhs = torch.randn(B, N, 1, device=dev, generator=gen)
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


left = torch.randn(B, N, S + 1, device=dev, generator=gen)
right = torch.randn(B, N, S + 1, device=dev, generator=gen)
out = torch.empty(B, N, N, device=dev)
ms = timed(lambda: _kernels.bilinear_dense(left, right, None, out=out.zero_()) if False else _kernels.bilinear_dense(left, right))
alg = 4.0 * B * (N * N + 2 * N * (S + 1))
print(json.dumps({"kernel": "k_bilinear_dense<float,8,8> (dense operator gradient, rank-33 outer product)", "batch": B,
                  "ms": ms, "algorithmic_GB": alg / 1e9, "GBps": alg / ms / 1e6, "frac_of_hbm_peak": alg / ms / 1e6 / peak}))
del left, right, out
torch.cuda.empty_cache()

K.requires_grad_(True)
d.requires_grad_(True)
rhs.requires_grad_(True)


def fwd_bwd():
    K.grad = d.grad = rhs.grad = None
    op = AddedDiagLinearOperator(DenseLinearOperator(K), DiagLinearOperator(d))
    iq, ld = op.inv_quad_logdet(rhs, logdet=True)
    (iq + ld).sum().backward()


def fwd():
    with torch.no_grad():
        op = AddedDiagLinearOperator(DenseLinearOperator(K), DiagLinearOperator(d))
        op.inv_quad_logdet(rhs, logdet=True)


with settings.num_trace_samples(S), settings.max_preconditioner_size(100):
    t_f = timed(fwd, 3)
    t_fb = timed(fwd_bwd, 3)
print(json.dumps({"call": "inv_quad_logdet forward + backward (operator, diagonal and rhs gradients), cold", "batch": B,
                  "forward_ms": t_f, "forward_backward_ms": t_fb, "backward_ms": t_fb - t_f,
                  "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}))

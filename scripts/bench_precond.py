"""Times the two products of the preconditioner apply z = (r - Q Q^T r)/s at config-2 shapes (B x 5000 x 100, 33 columns)
through each dense kernel (pinned with _lib.pin_dense_impl) and the CUDA-core Q^T r kernel.  Usage: python scripts/bench_precond.py [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linear_operator_b200 import _kernels, _lib  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N, k, C = 5000, 100, 33
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
Q = torch.randn(B, N, k, device=dev, generator=g) / N**0.5
r = torch.randn(B, N, C, device=dev, generator=g)
t = torch.randn(B, k, C, device=dev, generator=g)
alpha = -torch.ones(B, device=dev)
d = torch.full((B, 1), 2.0, device=dev)


def timeit(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ref = None
for impl in ["tc", "stream2", "simt"]:
    _lib.pin_dense_impl(impl)
    ms = timeit(lambda: _kernels.dense_matmul(Q, t, d=d, want_dots=True, E=r, alpha=alpha))
    z = _kernels.dense_matmul(Q, t, d=d, want_dots=True, E=r, alpha=alpha)[0]
    if ref is None:
        ref = (alpha.view(B, 1, 1).double() * (Q.double() @ t.double()) + 2.0 * r.double())
    err = ((z.double() - ref).abs().max() / ref.abs().max()).item()
    print(f"Q t + epilogue  impl={impl:8s} B={B}: {ms:7.3f} ms   err {err:.2e}")
_lib.pin_dense_impl(None)
ms = timeit(lambda: _kernels.tn_matmul(Q, r))
print(f"Q^T r (tn_matmul, CUDA cores)   B={B}: {ms:7.3f} ms")

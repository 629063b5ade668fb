"""Dense operator matmul Y = A X (+ d (.) X) at small batch counts (single-GP training shapes): this library's
dispatch against torch.matmul (cuBLAS SGEMM, what the reference runs) and against lob_gemm3x with split-K."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linear_operator_b200 import _kernels  # noqa: E402

dev = "cuda:0"


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


torch.backends.cuda.matmul.allow_tf32 = False
for B, N in [(1, 2000), (1, 5000), (1, 10000), (1, 20000), (4, 5000), (16, 5000), (64, 5000)]:
    g = torch.Generator(device=dev).manual_seed(0)
    A = torch.randn(B, N, N, device=dev, generator=g) / N**0.5
    X = torch.randn(B, N, 33, device=dev, generator=g)
    d = torch.rand(B, N, device=dev, generator=g)
    ref = A.double() @ X.double() + d.double().unsqueeze(-1) * X.double()
    t_ours = timed(lambda: _kernels.dense_matmul(A, X, d=d, want_dots=True))
    y = _kernels.dense_matmul(A, X, d=d)
    err = ((y.double() - ref).abs().max() / ref.abs().max()).item()
    t_cublas = timed(lambda: torch.addcmul(A @ X, d.unsqueeze(-1), X))
    bytes_ = 4.0 * B * N * N
    print(f"B={B:3d} N={N:6d}: ours {t_ours:8.1f} us ({bytes_ / t_ours * 1e-6:6.2f} TB/s, err {err:.1e}) | "
          f"torch.matmul + addcmul {t_cublas:8.1f} us ({bytes_ / t_cublas * 1e-6:6.2f} TB/s)")

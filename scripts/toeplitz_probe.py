"""One symmetric Toeplitz product at BASELINE config 4's shape (N = 2^20, 33 columns, batch 16): CUDA-event time; run it under
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum for the per-kernel breakdown."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from linear_operator_b200 import _kernels
dev = "cuda:0"
B, N, C = 16, 2**20, 33
j = torch.arange(N, device=dev, dtype=torch.float32)
col = torch.exp(-0.5 * (j[None, :] / 50.0) ** 2).repeat(B, 1)
X = torch.randn(B, N, C, device=dev)
d = torch.full((B, N), 0.5, device=dev)
fc = _kernels.toeplitz_embed_fft(col)
def timed(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for rnd in range(3):  # alternate the two forms: box-to-box and run-to-run noise is larger than their difference
    a = timed(lambda: _kernels.toeplitz_matmul(col, X, d, fc_cache=fc))
    b = timed(lambda: _kernels.toeplitz_matmul(col, X, d, fc_cache=fc, want_dots=True))
    print(f"toeplitz_matmul {a:.3f} ms | with fused <X, Y> partials {b:.3f} ms")

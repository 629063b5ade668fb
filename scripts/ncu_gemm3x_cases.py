"""One launch each of the three shapes lob_gemm3x serves on the path, for `ncu --set full -k regex:k_gemm3x`:
(1) W = R U, split-K (BASELINE config 5 at N = 2e6);  (2) x = (R - w U^T)/sigma, fused epilogue;  (3..5) the three mode
products of the Kronecker chain (100 x 100 x 100, batch 2, 33 columns)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linear_operator_b200 import _kernels  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(5)
B, N, r = 512, 2_000_000, 256
U = torch.randn(N, r, device=dev, generator=g) / 16
R = torch.randn(B, N, device=dev, generator=g)
w = torch.randn(B, r, device=dev, generator=g) / 100
sig = 0.5 + torch.rand(B, device=dev, generator=g)
_kernels.gemm3x(R.unsqueeze(0), U.unsqueeze(0))
_kernels.gemm3x(w.unsqueeze(0), U.unsqueeze(0), trans_b=True, row_alpha=(-1 / sig).unsqueeze(0), E=R.unsqueeze(0),
                row_beta=(1 / sig).unsqueeze(0))
del U, R
fs = [torch.randn(2, 100, 100, device=dev, generator=g) for _ in range(3)]
X = torch.randn(2, 100**3, 33, device=dev, generator=g)
_kernels.kron_matmul(fs, X)
torch.cuda.synchronize()

import sys, os
import torch
sys.path.insert(0, '/root/repo')
from linear_operator_b200 import _kernels
B, N, C = 8, 1024, 33
g = torch.Generator(device='cuda').manual_seed(5)
A = torch.randn(B, N, N, device='cuda', generator=g) / N**0.5
X = torch.randn(B, N, C, device='cuda', generator=g)
Y = _kernels.dense_matmul(A, X)
torch.cuda.synchronize()
print("ok", float(Y.abs().sum()))

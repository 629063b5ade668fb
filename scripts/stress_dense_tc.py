import sys, time
import torch
sys.path.insert(0, '/root/repo')
from linear_operator_b200 import _kernels
B, N, C = 48, 5000, 33
g = torch.Generator(device='cuda').manual_seed(5)
A = torch.randn(B, N, N, device='cuda', generator=g) / N**0.5
X = torch.randn(B, N, C, device='cuda', generator=g)
import os
os.environ["LOB_DISABLE_TC"] = "1"
ref = _kernels.dense_matmul(A, X)
del os.environ["LOB_DISABLE_TC"]
scale = ref.abs().max()
bad_tot = 0
t0 = time.time()
for it in range(12):
    Y = _kernels.dense_matmul(A, X)
    bad_tot += int((((Y - ref).abs() / scale) > 1e-4).sum())
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for it in range(5):
    Y = _kernels.dense_matmul(A, X)
e1.record(); torch.cuda.synchronize()
print(f"dbg={os.environ.get('LOB_TC_DBG','0')}: bad elements over 12 launches: {bad_tot}; {e0.elapsed_time(e1)/5:.3f} ms/launch")

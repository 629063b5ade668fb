import sys, os
import torch
sys.path.insert(0, '/root/repo')
from linear_operator_b200 import _kernels
B, N, C = 48, 5000, 33
g = torch.Generator(device='cuda').manual_seed(5)
A = torch.randn(B, N, N, device='cuda', generator=g) / N**0.5
X = torch.randn(B, N, C, device='cuda', generator=g)
os.environ["LOB_DISABLE_TC"] = "1"
ref = _kernels.dense_matmul(A, X)
del os.environ["LOB_DISABLE_TC"]
scale = ref.abs().max()
found = 0
for it in range(12):
    Y = _kernels.dense_matmul(A, X)
    bad = (((Y - ref).abs() / scale) > 1e-4).any(-1).nonzero()
    for (b0, r0) in bad[:6].tolist():
        a = torch.zeros(157 * 32, dtype=torch.float64, device='cuda'); a[:N] = A[b0, r0].double()
        x = torch.zeros(157 * 32, C, dtype=torch.float64, device='cuda'); x[:N] = X[b0].double()
        diff = (Y[b0, r0].double() - ref[b0, r0].double())
        contrib = torch.einsum('bk,bkc->bc', a.view(157, 32), x.view(157, 32, C))      # per k-block
        prefix = torch.cumsum(contrib, 0)                                               # sum of blocks 0..k0
        res_all = (diff.unsqueeze(0) + prefix).abs()                                    # (157, C)
        best_k = res_all.amax(-1).argmin().item()
        percol = res_all.min(0)
        print(f"row ({b0},{r0},%256={r0%256}) |diff| {diff.abs().max().item():.2e}: prefix-loss fit all-cols: k0={best_k} resid {res_all.amax(-1).min().item():.2e}; per-column best resid max {percol.values.max().item():.2e} k0s {sorted(set(percol.indices.tolist()))[:8]}")
        found += 1
    if found >= 8: break
print("done, bad rows analysed:", found)

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
os.environ["LOB_TC_DBG"] = "1024"
scale = ref.abs().max()
found = 0
for it in range(12):
    Y = _kernels.dense_matmul(A, X)
    bad = (((Y - ref).abs() / scale) > 1e-4).any(-1).nonzero()
    for (b0, r0) in bad[:5].tolist():
        m0 = (r0 // 256) * 256
        rows = torch.arange(m0, min(m0 + 256, N), device='cuda')
        Ab = torch.zeros(rows.numel(), 157 * 32, dtype=torch.float64, device='cuda'); Ab[:, :N] = A[b0, rows].double()
        x = torch.zeros(157 * 32, C, dtype=torch.float64, device='cuda'); x[:N] = X[b0].double()
        diff = (Y[b0, r0].double() - ref[b0, r0].double())
        Ab = Ab.view(-1, 157, 32); xb = x.view(157, 32, C)
        contrib = torch.einsum('rbk,bkc->rbc', Ab, xb)          # (R, 157, C)
        mine = contrib[r0 - m0]                                  # (157, C)
        cand = contrib - mine.unsqueeze(0)                       # row r' data used for block kb
        res = (cand - diff).abs().amax(-1)                       # (R, 157)
        i = int(res.argmin()); rp, kb = divmod(i, 157)
        print(f"row ({b0},{r0},%256={r0%256}) |diff| {diff.abs().max().item():.2e}; best other-row fit: resid {res.min().item():.2e} r'%256={rp} kb={kb}")
        found += 1
    if found >= 8: break

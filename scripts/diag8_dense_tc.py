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
    for (b0, r0) in bad[:5].tolist():
        a = torch.zeros(157 * 32, dtype=torch.float64, device='cuda'); a[:N] = A[b0, r0].double()
        x = torch.zeros(157 * 32, C, dtype=torch.float64, device='cuda'); x[:N] = X[b0].double()
        diff = (Y[b0, r0].double() - ref[b0, r0].double())
        ab = a.view(157, 32); xb = x.view(157, 32, C)
        res = []
        for dj in (2, 5, 1, 3):
            for kb in range(dj, 156):
                atoms = (ab[kb - dj] - ab[kb]).unsqueeze(-1) * xb[kb]     # (32, C): column c of block stale
                w = torch.linalg.lstsq(atoms.T, diff.unsqueeze(-1)).solution[:, 0]
                r = (atoms.T @ w - diff).abs().max().item()
                q = torch.minimum(w.abs(), (w - 1).abs()).max().item()   # distance of weights from {0,1}
                res.append((q, r, dj, kb, int((w > 0.5).sum())))
        res.sort()
        print(f"row ({b0},{r0},%256={r0%256}) |diff| {diff.abs().max().item():.2e}; best 0/1-weight fits (dist, resid, dj, kb, nstale): {[(round(q,3), float(f'{r:.1e}'), dj, kb, n) for q, r, dj, kb, n in res[:3]]}")
        found += 1
    if found >= 6: break

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
    for (b0, r0) in bad[:4].tolist():
        a = torch.zeros(157 * 32, dtype=torch.float64, device='cuda'); a[:N] = A[b0, r0].double()
        x = torch.zeros(157 * 32, C, dtype=torch.float64, device='cuda'); x[:N] = X[b0].double()
        diff = (Y[b0, r0].double() - ref[b0, r0].double())
        a4 = a.view(157, 4, 8); x4 = x.view(157, 4, 8, C)
        out = []
        for dj in (1, 2, 3, 5, 10):
            best = None
            for kb in range(dj, 157):
                # basis: k-step s of block kb computed with A of block kb-dj
                basis = torch.einsum('sk,skc->sc', a4[kb - dj] - a4[kb], x4[kb])   # (4, C)
                sol = torch.linalg.lstsq(basis.T, diff.unsqueeze(-1)).solution[:, 0]
                r = (basis.T @ sol - diff).abs().max().item()
                if best is None or r < best[0]: best = (r, kb, [round(v, 2) for v in sol.tolist()])
            out.append((dj, best))
        print(f"row ({b0},{r0},%256={r0%256}) |diff| {diff.abs().max().item():.2e}: " + "; ".join(f"dj={dj}: res {b[0]:.1e} kb={b[1]} w={b[2]}" for dj, b in out))
        found += 1
    if found >= 6: break

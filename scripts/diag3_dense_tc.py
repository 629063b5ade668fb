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
        ab = a.view(157, 32); xb = x.view(157, 32, C)
        # P[j, kb] = a_block_j @ x_block_kb  -> (157,157,C)
        P = torch.einsum('jk,bkc->jbc', ab, xb)
        true = torch.einsum('bk,bkc->bc', ab, xb)            # (157, C)
        cand = P - true.unsqueeze(0)                            # replace block kb's A by block j's
        res = (cand - diff).abs().amax(-1)                      # (157 j, 157 kb)
        m = res.min(); j, kb = divmod(int(res.argmin()), 157)
        # also hypothesis: block kb dropped (A = 0) or doubled
        drop = (-true - diff).abs().amax(-1); dbl = (true - diff).abs().amax(-1)
        print(f"launch {it} row ({b0},{r0}, row%256={r0%256}): |diff| {diff.abs().max().item():.3e}; best stale: res {m.item():.2e} kb={kb} j={j} (kb-j={kb-j}); best drop res {drop.min().item():.2e} kb={int(drop.argmin())}; best double res {dbl.min().item():.2e}")
        found += 1
    if found >= 10: break

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
        # hypothesis X-stale: block kb computed with X of block j, per k-step subsets (4 k-steps of 8)
        best = None
        true_steps = torch.einsum('bsk,bskc->bsc', ab.view(157, 4, 8), xb.view(157, 4, 8, C))  # (157,4,C)
        for dj in (1, 2, 3, 4, 5, 6):
            for kb in range(dj, 157):
                j = kb - dj
                stale_steps = torch.einsum('sk,skc->sc', ab[kb].view(4, 8), xb[j].view(4, 8, C))  # (4,C)
                delta = stale_steps - true_steps[kb]      # (4, C) per k-step contribution of the error
                for mask in range(1, 16):
                    sel = torch.tensor([(mask >> s) & 1 for s in range(4)], dtype=torch.float64, device='cuda')
                    cand = (delta * sel[:, None]).sum(0)
                    r = (cand - diff).abs().max().item()
                    if best is None or r < best[0]: best = (r, kb, dj, mask)
        print(f"launch {it} row ({b0},{r0}, row%256={r0%256}): |diff| {diff.abs().max().item():.3e}; best X-stale: res {best[0]:.2e} kb={best[1]} dj={best[2]} kstep-mask={best[3]:04b}")
        found += 1
    if found >= 6: break

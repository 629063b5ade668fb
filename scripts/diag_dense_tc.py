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
for it in range(4):
    Y = _kernels.dense_matmul(A, X)
    err = (Y - ref).abs() / scale
    bad = err > 1e-4
    idx = bad.nonzero()
    print(f"launch {it}: bad {idx.shape[0]} maxerr {err.max().item():.3e}")
    if idx.shape[0]:
        # group by (batch, 256-row cta)
        key = idx[:, 0] * 1000 + idx[:, 1] // 256
        uk, cnt = key.unique(return_counts=True)
        print("  ctas affected", uk.numel(), "of", B * 20, " counts per cta (first 10)", cnt[:10].tolist())
        k0 = uk[0].item(); b0, c0 = k0 // 1000, k0 % 1000
        sub = idx[(idx[:, 0] == b0) & (idx[:, 1] // 256 == c0)]
        rows = (sub[:, 1] - c0 * 256)
        print("  first cta: batch", b0, "cta", c0, "rows(min,max,n unique)", rows.min().item(), rows.max().item(), rows.unique().numel(), "cols unique", sub[:, 2].unique().tolist()[:40])
        r0 = sub[0, 1].item()
        print("  sample row", r0, "err per col", [f"{v:.1e}" for v in err[b0, r0].tolist()][:12])
        # is the error equal to a missing/doubled k-block contribution? compare with partial products
        a = A[b0, r0].double(); x = X[b0].double()
        diff = (Y[b0, r0].double() - ref[b0, r0].double())
        best = None
        for kb in range(157):
            part = a[kb*32:(kb+1)*32] @ x[kb*32:(kb+1)*32]
            for sgn, name in ((1, "doubled"), (-1, "missing")):
                r = (diff - sgn * part).abs().max().item()
                if best is None or r < best[0]: best = (r, kb, name)
        print("  best single-k-block explanation: residual", f"{best[0]:.2e}", "kb", best[1], best[2], " |diff|max", f"{diff.abs().max().item():.2e}")

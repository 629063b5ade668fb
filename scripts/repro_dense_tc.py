import sys
import torch
sys.path.insert(0, '/root/repo')
from linear_operator_b200 import _kernels

B, N, C = int(sys.argv[1]) if len(sys.argv) > 1 else 4, 5000, 33
g = torch.Generator(device='cuda').manual_seed(11)
A = torch.randn(B, N, N, device='cuda', generator=g) / N**0.5
X = torch.randn(B, N, C, device='cuda', generator=g)
d = torch.rand(B, N, device='cuda', generator=g) + 0.1
ref = (A.double() @ X.double() + d.double().unsqueeze(-1) * X.double())
scale = ref.abs().max()
outs = []
for i in range(8):
    Y, dots, _ = _kernels.dense_matmul(A, X, d=d, want_dots=True)
    outs.append((Y, dots))
torch.cuda.synchronize()
for i, (Y, dd) in enumerate(outs):
    err = ((Y.double() - ref).abs() / scale)
    bad = err > 1e-4
    nb = int(bad.sum())
    msg = f"{i}: maxerr {err.max().item():.3e} bad {nb} Yeq0 {torch.equal(Y, outs[0][0])} dotseq0 {torch.equal(dd, outs[0][1])}"
    if nb:
        idx = bad.nonzero()
        bs = sorted(set(idx[:, 0].tolist()))
        rows = idx[:, 1]
        cols = sorted(set(idx[:, 2].tolist()))
        tiles = sorted(set((rows // 128).tolist()))
        msg += f" batches {bs} row-tiles(128) {tiles[:12]}{'...' if len(tiles)>12 else ''} ncols {len(cols)} cols {cols[:8]} rows/tile {nb/ max(1,len(tiles)) / max(1,len(cols)):.1f}"
    print(msg)

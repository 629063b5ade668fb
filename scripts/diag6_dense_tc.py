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
        atoms = torch.einsum('ak,akc->ac', a.view(628, 8), x.view(628, 8, C))   # contribution of each 8-wide k-step
        res = diff.clone(); picked = []
        for step in range(6):
            # best atom with coefficient -1 (dropped) or +1 (doubled)
            rm = (res.unsqueeze(0) + atoms).abs().amax(-1)   # dropped: diff = -atom -> res + atom small
            rp = (res.unsqueeze(0) - atoms).abs().amax(-1)
            im, ip = int(rm.argmin()), int(rp.argmin())
            if rm[im] <= rp[ip]:
                res = res + atoms[im]; picked.append((im // 4, im % 4, 'drop'))
            else:
                res = res - atoms[ip]; picked.append((ip // 4, ip % 4, 'dbl'))
            if res.abs().max() < 1e-3 * max(diff.abs().max().item(), 1e-9): break
        print(f"row ({b0},{r0},%256={r0%256}) |diff| {diff.abs().max().item():.2e} -> residual {res.abs().max().item():.2e} after {len(picked)} atoms {picked}")
        found += 1
    if found >= 8: break

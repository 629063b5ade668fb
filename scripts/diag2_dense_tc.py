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
hist = torch.zeros(256, dtype=torch.long, device='cuda')
ncta = 0
for it in range(12):
    Y = _kernels.dense_matmul(A, X)
    bad = (((Y - ref).abs() / scale) > 1e-4).any(-1)  # (B, N) rows
    idx = bad.nonzero()
    if idx.numel():
        hist += torch.bincount(idx[:, 1] % 256, minlength=256)
        ncta += (idx[:, 0] * 1000 + idx[:, 1] // 256).unique().numel()
h = hist.cpu().tolist()
print("bad rows total", sum(h), "ctas", ncta)
print("by warp-quarter (32-row groups):", [sum(h[i*32:(i+1)*32]) for i in range(8)])
print("rows<32 detail:", h[:32])
# for one bad row, find which k-blocks are wrong: compare against products with A rows zeroed per block
Y = _kernels.dense_matmul(A, X)
bad = (((Y - ref).abs() / scale) > 1e-4).any(-1).nonzero()
if bad.numel():
    b0, r0 = bad[0].tolist()
    a = A[b0, r0].double(); x = X[b0].double(); diff = (Y[b0, r0].double() - ref[b0, r0].double())
    # least squares: diff ~ sum_kb w_kb * (a_kb @ x_kb): solve for w (157 unknowns, 33 equations -> underdetermined); instead test
    # hypothesis "block kb used stale A from block kb-2": diff = a[kb-2] @ x[kb] - a[kb] @ x[kb]
    best = []
    for kb in range(2, 157):
        xs = x[kb*32:(kb+1)*32]
        n = xs.shape[0]
        cand = a[(kb-2)*32:(kb-2)*32+n] @ xs - a[kb*32:kb*32+n] @ xs
        best.append(((diff - cand).abs().max().item(), kb))
    best.sort()
    print("row", b0, r0, "|diff|max", diff.abs().max().item(), "best stale(kb-2) hypotheses:", best[:3])

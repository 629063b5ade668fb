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
        y = Y[b0, r0]
        d_same_batch = (ref[b0] - y).abs().amax(-1)
        j = int(d_same_batch.argmin())
        d_all = (ref.reshape(-1, C) - y).abs().amax(-1)
        ja = int(d_all.argmin())
        diff = (y - ref[b0, r0])
        print(f"row ({b0},{r0},%256={r0%256}): |diff| {diff.abs().max().item():.2e}; nearest ref row in batch: {j} dist {d_same_batch[j].item():.2e}; nearest anywhere: (b={ja//N}, r={ja%N}) dist {d_all[ja].item():.2e}")
        print("    y  :", [f"{v:+.3f}" for v in y[:10].tolist()])
        print("    ref:", [f"{v:+.3f}" for v in ref[b0, r0, :10].tolist()])
        print("    dif:", [f"{v:+.3f}" for v in diff[:10].tolist()], " ratio y/ref:", [f"{(a/b):.3f}" for a, b in zip(y[:6].tolist(), ref[b0, r0, :6].tolist())])
        found += 1
    if found >= 6: break

"""BASELINE configs[0]: Dense + AddedDiag, N = 512, batch 1, fp64, 16 probes, CG path forced (max_cholesky_size(0))."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linear_operator_b200 import settings
from linear_operator_b200.operators import AddedDiagLinearOperator, DenseLinearOperator, DiagLinearOperator
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1234)
N = 512
W = torch.randn(1, N, 256, device=dev, generator=g, dtype=torch.float64)
sc = torch.logspace(0, -1.5, 256, device=dev, dtype=torch.float64); W = W * sc / sc.norm()
K = W @ W.mT
d = torch.full((1, N), 0.5, device=dev, dtype=torch.float64)
rhs = torch.randn(1, N, 1, device=dev, generator=g, dtype=torch.float64)
def step():
    op = AddedDiagLinearOperator(DenseLinearOperator(K), DiagLinearOperator(d))
    return op.inv_quad_logdet(rhs, logdet=True)
with settings.max_cholesky_size(0), settings.num_trace_samples(16):
    for _ in range(5): iq, ld = step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(50): iq, ld = step()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 50
exact = torch.logdet(K[0] + torch.diag(d[0]))
print(f"cfg1: {1/dt:.1f} calls/s ({dt*1e3:.2f} ms per cold call); logdet {ld.item():.6f} vs exact {exact.item():.6f}")

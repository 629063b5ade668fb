"""Times the pieces of BASELINE config 3 (Kronecker 100^3, batch 8, 33 columns): mode-product matmul, pivoted Cholesky."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linear_operator_b200 import _kernels, settings
from linear_operator_b200.operators import AddedDiagLinearOperator, DiagLinearOperator, KroneckerProductLinearOperator
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(3)
fs = []
for _ in range(3):
    G = torch.randn(8, 100, 100, device=dev, generator=g)
    fs.append(G @ G.mT / 100 + 0.1 * torch.eye(100, device=dev))
N = 100**3
kron = KroneckerProductLinearOperator(*fs)
x = torch.randn(8, N, 33, device=dev, generator=g)
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print("kron matmul (8 x 1e6 x 33):", timeit(lambda: kron._matmul(x)), "ms  (algorithmic 2.11 GB, 158 GFLOP)")
print("pivoted cholesky rank 100:", timeit(lambda: kron._pivoted_cholesky(100, 1e-3), 1), "ms")
op = AddedDiagLinearOperator(kron, DiagLinearOperator(torch.full((8, N), 0.5, device=dev)))
with settings.max_preconditioner_size(100):
    pc = op._preconditioner()[0]
    r = torch.randn(8, N, 33, device=dev, generator=g)
    print("preconditioner apply:", timeit(lambda: pc(r)), "ms")

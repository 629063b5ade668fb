"""BASELINE configs[0]: Dense + AddedDiag, N = 512, batch 1, fp64, 16 probes, CG path forced (max_cholesky_size(0)).

    python scripts/bench_cfg1.py [ours|reference-gpu|reference-cpu]

Prints one JSON line; `reference-*` runs the UNMODIFIED reference from baseline/_ref on the named device."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
impl = sys.argv[1] if len(sys.argv) > 1 else "ours"
if impl == "ours":
    from linear_operator_b200 import settings
    from linear_operator_b200.operators import AddedDiagLinearOperator, DenseLinearOperator, DiagLinearOperator
else:
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    from linear_operator import settings
    from linear_operator.operators import AddedDiagLinearOperator, DenseLinearOperator, DiagLinearOperator
dev = torch.device("cpu" if impl == "reference-cpu" else "cuda:0")
if dev.type == "cpu":
    torch.set_num_threads(len(os.sched_getaffinity(0)))
g = torch.Generator(device=dev).manual_seed(1234)
N = 512
W = torch.randn(1, N, 256, device=dev, generator=g, dtype=torch.float64)
sc = torch.logspace(0, -1.5, 256, device=dev, dtype=torch.float64)
W = W * sc / sc.norm()
K = W @ W.mT
d = torch.full((1, N), 0.5, device=dev, dtype=torch.float64)
rhs = torch.randn(1, N, 1, device=dev, generator=g, dtype=torch.float64)


def step():
    op = AddedDiagLinearOperator(DenseLinearOperator(K), DiagLinearOperator(d))
    return op.inv_quad_logdet(rhs, logdet=True)


def sync():
    if dev.type == "cuda":
        torch.cuda.synchronize()


def measure(reps=50):
    for _ in range(5):
        step()
    sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = step()
    sync()
    return out, (time.perf_counter() - t0) / reps


extra = {}
with settings.max_cholesky_size(0), settings.num_trace_samples(16), torch.no_grad():
    (iq, ld), dt = measure()
    if impl == "ours":
        with settings.cuda_graphs(False):
            _, dt_eager = measure()
        extra = {"ms_per_call_without_cuda_graph": dt_eager * 1e3, "calls_per_s_without_cuda_graph": 1 / dt_eager}
        # the solver alone (21 iterations, 17 columns), graph replay vs launch by launch
        from linear_operator_b200.utils import linear_cg

        op = AddedDiagLinearOperator(DenseLinearOperator(K), DiagLinearOperator(d))
        rhs17 = torch.randn(1, N, 17, device=dev, generator=g, dtype=torch.float64)

        def solve():
            return linear_cg(op._matmul_closure(), rhs17, n_tridiag=16, _skip_initial_matmul=True)

        for flag in (True, False):
            with settings.cuda_graphs(flag):
                for _ in range(5):
                    solve()
                sync()
                t0 = time.perf_counter()
                for _ in range(50):
                    solve()
                sync()
                extra["linear_cg_ms_graph" if flag else "linear_cg_ms_eager"] = (time.perf_counter() - t0) / 50 * 1e3
exact = torch.logdet(K[0] + torch.diag(d[0]))
print(json.dumps({"config": "BASELINE configs[0]: Dense+AddedDiag N=512 batch 1 fp64 16 probes, cold calls", "impl": impl,
                  "device": str(dev), "calls_per_s": 1 / dt, "ms_per_call": dt * 1e3, "cg_iters_per_s": 21 / dt,
                  "logdet": float(ld), "logdet_exact": float(exact),
                  "threads": torch.get_num_threads() if dev.type == "cpu" else None, **extra}))

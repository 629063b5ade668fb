"""Per-entry-point GPU time of one cold inv_quad_logdet at BASELINE config 2, measured in situ with CUDA events on the
launching stream (no profiler: kernels overlap and cache exactly as in bench.py).  Every ``_kernels`` wrapper and every
``lob_cg_*`` call is bracketed by an event pair; `unaccounted` = step time - sum of the brackets = torch glue kernels
(randn, cat, isnan, memsets) + device idle time behind host work / synchronisations.

    python scripts/phase_profile.py [batch] [dense|toeplitz|kron] [pinned dense kernel, e.g. stream2: A/B on one box]
"""
import collections
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linear_operator_b200 as lo  # noqa: E402
from linear_operator_b200 import _kernels, _lib, settings  # noqa: E402
from linear_operator_b200.operators import (AddedDiagLinearOperator, DenseLinearOperator, DiagLinearOperator,  # noqa: E402
                                            KroneckerProductLinearOperator, ToeplitzLinearOperator)

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
KIND = sys.argv[2] if len(sys.argv) > 2 else "dense"
S = 32
if len(sys.argv) > 3:
    _lib.pin_dense_impl(sys.argv[3])
gen = torch.Generator(device=dev).manual_seed(1234)
if KIND == "dense":
    N = 5000
    sc = torch.logspace(0, -1.5, 256, device=dev)
    sc = sc / sc.norm()
    K = torch.empty(B, N, N, device=dev)
    for s0 in range(0, B, 32):
        W = torch.randn(min(32, B - s0), N, 256, device=dev, generator=gen) * sc
        torch.bmm(W, W.mT, out=K[s0:s0 + W.shape[0]])
    del W
    make_base = lambda: DenseLinearOperator(K)  # noqa: E731
elif KIND == "toeplitz":
    N = 2**20
    j = torch.arange(N, device=dev, dtype=torch.float32)
    ls = 50.0 * (1 + torch.arange(B, device=dev, dtype=torch.float32) / 64)
    col = torch.exp(-0.5 * (j[None, :] / ls[:, None]) ** 2)
    make_base = lambda: ToeplitzLinearOperator(col)  # noqa: E731
else:
    N = 100**3
    fs = []
    for _ in range(3):
        G = torch.randn(B, 100, 100, device=dev, generator=gen)
        fs.append(G @ G.mT / 100 + 0.1 * torch.eye(100, device=dev))
    make_base = lambda: KroneckerProductLinearOperator(*fs)  # noqa: E731
d = torch.full((B, N), 0.5, device=dev)
rhs = torch.randn(B, N, 1, device=dev, generator=gen)

records = collections.defaultdict(list)
ENABLED = [False]


def bracket(name, fn):
    def wrapped(*a, **k):
        if not ENABLED[0]:
            return fn(*a, **k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*a, **k)
        e1.record()
        records[name].append((e0, e1))
        return out

    return wrapped


# wrappers called from inside another bracketed wrapper would be counted twice
NESTED = {"gemm3x"} if len(sys.argv) > 2 and sys.argv[2] == "kron" else set()
for name in dir(_kernels):
    obj = getattr(_kernels, name)
    if name in NESTED:
        continue
    if callable(obj) and not name.startswith("_") and getattr(obj, "__module__", "") == _kernels.__name__ \
            and not isinstance(obj, type):
        setattr(_kernels, name, bracket(name, obj))
lib = _lib.load()
for name in ("lob_cg_setup", "lob_cg_residual_init", "lob_cg_direction_init", "lob_cg_step_xr", "lob_cg_step_p",
             "lob_cg_finish", "lob_cg_step_fused"):
    if hasattr(lib, name):
        fn = getattr(lib, name)
        w = bracket(name, fn)
        setattr(lib, name, w)


def step():
    op = AddedDiagLinearOperator(make_base(), DiagLinearOperator(d))
    return op.inv_quad_logdet(rhs, logdet=True)


RANK = int(os.environ.get("PRECOND_RANK", "100"))  # BASELINE config 2 names rank 100; 15 is the reference's default
with settings.num_trace_samples(S), settings.max_preconditioner_size(RANK):
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    ENABLED[0] = True
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
total = e0.elapsed_time(e1) / reps
rows = []
acc = 0.0
for name, evs in records.items():
    ms = sum(a.elapsed_time(b) for a, b in evs) / reps
    rows.append((ms, name, len(evs) // reps))
    acc += ms
rows.sort(reverse=True)
print(f"cold inv_quad_logdet, {KIND}, N = {N}, batch {B}, preconditioner rank {RANK}{', dense kernel pinned to ' + sys.argv[3] if len(sys.argv) > 3 else ''}: "
      f"{total:.1f} ms per call")
for ms, name, n in rows:
    print(f"  {name:28s} {n:4d} calls  {ms:8.2f} ms  {100 * ms / total:5.1f} %")
print(f"  {'unaccounted (glue + idle)':28s}             {total - acc:8.2f} ms  {100 * (total - acc) / total:5.1f} %")
print(json.dumps({"kind": KIND, "n": N, "batch": B, "ms_per_call": total, "accounted_ms": acc,
                  "rows": [{"name": n, "calls": c, "ms": m} for m, n, c in rows]}))

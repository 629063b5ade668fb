"""BASELINE config 5's two tall products at one GPU's shard (batch 512, N = 10^7, r = 256) through lob_gemm3x:
time, fraction of the HBM bound, and error against an fp64 product of the same fp32 inputs, for several split-K counts."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linear_operator_b200 import _kernels  # noqa: E402

dev = torch.device("cuda:0")
B, N, r = int(os.environ.get("BATCH", 512)), int(float(os.environ.get("NLONG", 1e7))), 256
g = torch.Generator(device=dev).manual_seed(5)
U = torch.randn(N, r, device=dev, generator=g) / 16
R = torch.randn(B, N, device=dev, generator=g)
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return out, e0.elapsed_time(e1) / reps


Ud = U.double()
w_ref = R[:16].double() @ Ud
for splits in (37, 1024, 0):
    w, ms = timed(lambda: _kernels.gemm3x(R.unsqueeze(0), U.unsqueeze(0), splits=splits))
    err = ((w[0, :16].double() - w_ref).norm() / w_ref.norm()).item()
    alg = 4.0 * (B * N + N * r + B * r)
    print(json.dumps({"product": "W = R U  (B x N)(N x r), split-K", "splits": splits, "ms": ms, "GBps": alg / ms / 1e6,
                      "frac_of_hbm_peak": alg / ms / 1e6 / peak, "tflops_fp32_equiv": 2.0 * B * N * r / ms / 1e9,
                      "relerr_vs_fp64": err}), flush=True)
wc = torch.randn(B, r, device=dev, generator=g) / 100
sig = 0.5 + torch.rand(B, device=dev, generator=g)
x, ms = timed(lambda: _kernels.gemm3x(wc.unsqueeze(0), U.unsqueeze(0), trans_b=True, row_alpha=(-1 / sig).unsqueeze(0),
                                      E=R.unsqueeze(0), row_beta=(1 / sig).unsqueeze(0)))
x_ref = (R[:16].double() - wc[:16].double() @ Ud.mT) / sig[:16].double().unsqueeze(-1)
alg = 4.0 * (2 * B * N + N * r)
print(json.dumps({"product": "x = (R - w U^T) / sigma  (B x r)(r x N), fused epilogue", "ms": ms, "GBps": alg / ms / 1e6,
                  "frac_of_hbm_peak": alg / ms / 1e6 / peak, "tflops_fp32_equiv": 2.0 * B * N * r / ms / 1e9,
                  "relerr_vs_fp64": ((x[0, :16].double() - x_ref).norm() / x_ref.norm()).item()}), flush=True)
G_ref = Ud.mT @ Ud
for splits in (1024, 4096):
    G, ms = timed(lambda: _kernels.gemm3x(U.unsqueeze(0), U.unsqueeze(0), trans_a=True, out_dtype=torch.float64, splits=splits))
    print(json.dumps({"product": "G = U^T U (r x r), split-K, fp64 output", "splits": splits, "ms": ms,
                      "relerr_vs_fp64": ((G[0] - G_ref).norm() / G_ref.norm()).item(),
                      "max_rel_diag_err": ((G[0].diagonal() - G_ref.diagonal()).abs() / G_ref.diagonal()).max().item()}), flush=True)
G2, ms = timed(lambda: _kernels.tn_matmul(U.unsqueeze(0), U.unsqueeze(0), out_dtype=torch.float64))
print(json.dumps({"product": "G = U^T U on the CUDA cores (k_matmul_tn<float,double>)", "ms": ms,
                  "relerr_vs_fp64": ((G2[0] - G_ref).norm() / G_ref.norm()).item()}), flush=True)
del Ud
w_t, ms_t = timed(lambda: R @ U)
print(json.dumps({"product": "torch.matmul(R, U) (cuBLAS fp32) for comparison", "ms": ms_t,
                  "relerr_vs_fp64": ((w_t[:16].double() - w_ref).norm() / w_ref.norm()).item()}), flush=True)

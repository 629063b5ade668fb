#!/usr/bin/env python3
"""bench.py -- headline benchmark of the batched Krylov hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): AddedDiagLinearOperator(DenseLinearOperator(K), DiagLinearOperator(0.5)),
N=5000, batch=1024 per GPU, fp32, 32 probe vectors, pivoted-Cholesky preconditioner of rank 100.  One "step" is one
cold ``op.inv_quad_logdet(rhs, logdet=True)`` forward on a fresh operator object (preconditioner rebuilt, as GP training
does every step).  Synthetic PSD inputs per SURVEY.md section 8d: K = W W^T, W = randn(B,N,256) * sc / |sc|,
sc = logspace(0,-1.5,256).

Prints ONE JSON line (rank 0).  `value` = calls/s with the operator resident in HBM; `e2e` = the same call fed from
pinned HOST memory (H2D of the operator + rhs and D2H of the results inside the timed region); `roofline` is for the
dominant kernel (the dense operator matmul); `cpu_baseline` is the numpy oracle port of the reference
timed on the host cores on a bounded batch slice.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CFG = dict(N=5000, B=1024, S=32, rank=100, wrank=256, diag=0.5, dtype="f32")
METRIC = "inv_quad_logdet calls/sec (batch-1024 N=5000 AddedDiag(Dense), fp32, 32 probes, rank-100 pivChol precond, cold)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=CFG["B"], help="per-GPU batch (default: the BASELINE config)")
    ap.add_argument("--n", type=int, default=CFG["N"])
    ap.add_argument("--cpu-sample-batch", type=int, default=16)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference (numpy, all host threads) on a bounded slice of the same workload
# ----------------------------------------------------------------------------------------------------------------
def cpu_inputs(bs, n, seed=1234):
    import numpy as np

    rng = np.random.default_rng(seed)
    W = rng.standard_normal((bs, n, CFG["wrank"]), dtype=np.float32)
    sc = np.logspace(0, -1.5, CFG["wrank"]).astype(np.float32)
    W = W * sc / np.linalg.norm(sc)
    K = W @ W.transpose(0, 2, 1)
    d = np.full((bs, n), CFG["diag"], np.float32)
    rhs = rng.standard_normal((bs, n, 1), dtype=np.float32)
    eps_root = rng.standard_normal((bs, CFG["rank"], CFG["S"]), dtype=np.float32)
    eps_diag = rng.standard_normal((CFG["S"], bs, n), dtype=np.float32)
    return K, d, rhs, eps_root, eps_diag


def cpu_step(inp):
    from oracle import krylov_oracle as ko

    K, d, rhs, eps_root, eps_diag = inp
    return ko.dense_added_diag_inv_quad_logdet(K, d, rhs, base_samples=(eps_root, eps_diag),
                                               precond_rank=CFG["rank"], min_precond_size=2000)


def time_cpu(bs, n, steps, warmup):
    inp = cpu_inputs(bs, n)
    for _ in range(warmup):
        cpu_step(inp)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step(inp)
    dt_step = (time.perf_counter() - t0) / steps
    return dt_step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    bs = args.cpu_sample_batch
    dt_step = time_cpu(bs, args.n, args.steps, args.warmup)
    calls_per_s = (bs / args.batch) / dt_step  # batch elements are independent: scale the slice to the full batch
    cores = os.cpu_count()
    sample = f"batch slice {bs} of {args.batch} (N={args.n}, same generator), scaled linearly to the full batch"
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": calls_per_s,
        "unit": "calls/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": dt_step * 1e3 * (args.batch / bs),
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": calls_per_s, "unit": "calls/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": calls_per_s, "unit": "calls/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "numpy oracle port of the reference's CPU path (the reference is pure Python/PyTorch and does not "
                "travel to the GPU box); all host threads through the BLAS",
    }
    emit(line)


def workload_config(args, world):
    return {
        "workload": f"BASELINE configs[1]: Dense+AddedDiag N={args.n}, batch={args.batch} per GPU, fp32, "
                    f"{CFG['S']} probes, pivoted-Cholesky precond rank={CFG['rank']}, 21 CG iterations (defaults)",
        "global_batch": args.batch * world,
        "call": "cold (fresh operator object per step: pivoted Cholesky + preconditioner factor + mBCG + SLQ)",
        "l2": "inputs (102 GB operator) are far larger than L2; no explicit flush",
        "parallelism": f"batch sharded over {world} GPU(s), one all_gather of the (inv_quad, logdet) results",
    }


# ----------------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import linear_operator_b200 as lo
    from linear_operator_b200 import _kernels, _lib, settings
    from linear_operator_b200.operators import AddedDiagLinearOperator, DenseLinearOperator, DiagLinearOperator

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: linear_operator_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    B, N, S = args.batch, args.n, CFG["S"]
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    sc = torch.logspace(0, -1.5, CFG["wrank"], device=dev)
    sc = sc / sc.norm()
    K = torch.empty(B, N, N, device=dev)
    chunk = 32
    for s in range(0, B, chunk):  # synthetic data generation (not timed)
        e = min(s + chunk, B)
        W = torch.randn(e - s, N, CFG["wrank"], device=dev, generator=gen) * sc
        torch.bmm(W, W.mT, out=K[s:e])
    del W
    d = torch.full((B, N), CFG["diag"], device=dev)
    rhs = torch.randn(B, N, 1, device=dev, generator=gen)
    torch.cuda.synchronize()

    ctx = [settings.num_trace_samples(S), settings.max_preconditioner_size(CFG["rank"])]
    for c in ctx:
        c.__enter__()

    def step(Kd, dd, rr):
        op = AddedDiagLinearOperator(DenseLinearOperator(Kd), DiagLinearOperator(dd))
        return op.inv_quad_logdet(rr, logdet=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    torch.manual_seed(4321 + rank)
    for _ in range(args.warmup):
        iq, ld = step(K, d, rhs)
    barrier()

    # ---- timed region: K steps, CUDA events, matmul kernel timed separately on the launching stream ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    _kernels.PROFILE_MATMUL = []
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        iq, ld = step(K, d, rhs)
    ev1.record()
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = _lib.launch_count() - launches0
    mm_events = _kernels.PROFILE_MATMUL
    _kernels.PROFILE_MATMUL = None
    clocks = sampler.stop() if rank == 0 else None
    mm_ms = [a.elapsed_time(b) for a, b, big in mm_events if big]

    # the single collective of the multi-GPU path: all_gather of the per-rank results (product code)
    from linear_operator_b200.distributed import gather_results

    iq_all, ld_all = gather_results(iq, ld, B * world)
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())

    ms_per_step = elapsed_ms / args.steps
    value = world * 1e3 / ms_per_step

    # ---- end to end: operator + rhs start in pinned HOST memory every step, results read back ----
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, torch, dist, world, dev, K, d, rhs, step, barrier)

    # ---- roofline of the dominant kernel ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    C = S + 1
    alg_bytes = 4.0 * B * (N * N + 2 * N * C)  # operator once + X read + Y write
    # DRAM traffic of the dominant kernel from the committed ncu capture (per batch element, scaled to this batch)
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r1_dense_stream2_traffic.json")))
        if int(tr["n"]) == N and int(tr["columns"]) == C:
            traffic = float(tr["dram_bytes_per_batch_element"]) * B
    except Exception:
        pass
    roof = None
    if mm_ms:
        avg_ms = sum(mm_ms) / len(mm_ms)
        achieved = alg_bytes / (avg_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "dense operator matmul Y = A X + d.X with fused <p,Ap> partials (k_split_x2 + k_dense_stream2)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                "traffic": traffic, "traffic_source": "ncu dram__bytes_read+write of k_dense_stream2 at batch 256, scaled "
                "per batch element (profiles/r1_dense_stream2_traffic.json)" if traffic else None,
                "avg_launch_ms": avg_ms, "launches_timed": len(mm_ms),
                "share_of_step": sum(mm_ms) / elapsed_ms if world == 1 else None,
                "algorithmic_bytes_per_launch": alg_bytes}

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        bs = args.cpu_sample_batch
        dt_step = time_cpu(bs, N, 1, 1)
        cpu_base = {"value": (bs / B) / dt_step, "unit": "calls/s", "cores": os.cpu_count(), "kind": "port",
                    "sample": f"batch slice {bs} of {B} (N={N}), 1 warm-up + 1 timed call of the numpy oracle, scaled "
                              "linearly to the full batch"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "calls/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
            "cg_iters_per_s": value * 21, "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e,
            "roofline": roof, "cpu_baseline": cpu_base,
            "result_check": {"inv_quad_mean": float(iq_all.mean()), "logdet_mean": float(ld_all.mean()),
                             "gathered": int(iq_all.numel())},
        }
        emit(line)
    for c in reversed(ctx):
        c.__exit__(None, None, None)
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, torch, dist, world, dev, K, d, rhs, step, barrier):
    """Same call fed from pinned host memory.  The host buffer holds as much of the operator as the box allows (the
    full 102 GB when it fits in a fraction of free RAM per rank); when it is smaller, the copy cycles through it so
    that the FULL operator byte count crosses PCIe every step (contents repeat -- the data is synthetic anyway)."""
    B, N = K.shape[0], K.shape[-1]
    per_elt = N * N * 4
    try:
        avail = 0
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable"):
                avail = int(line.split()[1]) * 1024
        budget = int(avail * 0.45 / max(world, 1))
    except Exception:
        budget = 8 << 30
    n_host = int(max(1, min(B, budget // per_elt)))
    try:
        Kh = torch.empty(n_host, N, N, pin_memory=True)
    except Exception:
        n_host = max(1, min(n_host, 16))
        Kh = torch.empty(n_host, N, N, pin_memory=True)
    Kh.copy_(K[:n_host])
    dh = d.cpu().pin_memory()
    rh = rhs.cpu().pin_memory()
    out_h = torch.empty(2, B, pin_memory=True)
    Kd = K  # reuse the device allocation as the destination of the per-step upload
    steps = max(1, min(args.steps, 2))
    # Upload and compute are pipelined over batch chunks (batch elements are independent): all H2D copies are queued
    # on a copy stream, the compute stream waits for chunk c's event, runs the public API call on that chunk and queues
    # the D2H of its results.  PCIe is the bottleneck (102 GB per step), the solver hides behind it.
    nchunk = 8 if B % 8 == 0 and B >= 64 else 1
    cb = B // nchunk
    copy_stream = torch.cuda.Stream(device=dev)
    dd_dev = torch.empty_like(d)
    rr_dev = torch.empty_like(rhs)

    def one():
        cur = torch.cuda.current_stream(dev)
        copy_stream.wait_stream(cur)
        events = []
        with torch.cuda.stream(copy_stream):
            for c in range(nchunk):
                lo_, hi_ = c * cb, (c + 1) * cb
                src0 = lo_ % max(1, n_host - cb + 1) if n_host >= cb else 0
                if n_host >= cb:
                    Kd[lo_:hi_].copy_(Kh[src0:src0 + cb], non_blocking=True)
                else:
                    for s0 in range(lo_, hi_, n_host):
                        e0 = min(s0 + n_host, hi_)
                        Kd[s0:e0].copy_(Kh[: e0 - s0], non_blocking=True)
                dd_dev[lo_:hi_].copy_(dh[lo_:hi_], non_blocking=True)
                rr_dev[lo_:hi_].copy_(rh[lo_:hi_], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                events.append(ev)
        for c in range(nchunk):
            lo_, hi_ = c * cb, (c + 1) * cb
            cur.wait_event(events[c])
            iq, ld = step(Kd[lo_:hi_], dd_dev[lo_:hi_], rr_dev[lo_:hi_])
            out_h[0, lo_:hi_].copy_(iq, non_blocking=True)
            out_h[1, lo_:hi_].copy_(ld, non_blocking=True)
        cur.synchronize()

    one()  # warm-up
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(steps):
        one()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms /= steps
    h2d = B * per_elt + dh.numel() * 4 + rh.numel() * 4
    return {"value": world * 1e3 / ms, "unit": "calls/s", "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(out_h.numel() * 4), "ms_per_step": ms, "steps": steps,
            "host_buffer_batch_elements": n_host,
            "chunks": nchunk,
            "note": "upload (copy stream) overlapped with compute over batch chunks; PCIe-bound"}


_JSON_FD = None


def emit(line: dict):
    """The ONE JSON line of the contract, written to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    args = parse()
    # Libraries print to stdout on their own (NCCL writes its version banner there when NCCL_DEBUG is set on the box):
    # keep fd 1 clean for the JSON line by pointing it at stderr for the duration of the run.
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

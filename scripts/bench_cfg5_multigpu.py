"""BASELINE configs[4]: LowRankRootAddedDiag, N = 10^7, rank 256, batch 4096 sharded over the GPUs of one box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/bench_cfg5_multigpu.py [--batch 4096] [--n 10000000] [--steps 5]

Rank g owns the contiguous batch slice shard_bounds(batch, g, world) (512 problems at 8 GPUs): the shared root U (N x 256,
10.2 GB) is replicated, the right-hand sides are generated per rank (no scatter), every rank runs the Woodbury path
(csrc/gemm3x.cu products + capacitance solve) on its slice and ONE all-gather assembles (inv_quad, logdet).  Prints one JSON
line on rank 0: calls/s of the WHOLE batch, max over ranks of the CUDA-event time."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linear_operator_b200.distributed import gather_results, shard_bounds  # noqa: E402
from linear_operator_b200.operators import ConstantDiagLinearOperator, LowRankRootLinearOperator  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--n", type=int, default=10**7)
ap.add_argument("--rank", type=int, default=256)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=2)
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
s0, s1 = shard_bounds(args.batch, rank, world)
nb = s1 - s0
gU = torch.Generator(device=dev).manual_seed(5)          # the same root on every rank
U = torch.randn(args.n, args.rank, device=dev, generator=gU) / 16
gR = torch.Generator(device=dev).manual_seed(100 + rank)  # this rank's right-hand sides
sig = (0.5 * (1 + torch.arange(s0, s1, device=dev, dtype=torch.float32) / args.batch)).reshape(nb, 1)
rhs = torch.randn(nb, args.n, 1, device=dev, generator=gR)


def step():
    op = LowRankRootLinearOperator(U) + ConstantDiagLinearOperator(sig, diag_shape=args.n)
    iq, ld = op.inv_quad_logdet(rhs, logdet=True)
    return gather_results(iq, ld, args.batch)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for _ in range(args.warmup):
    iq, ld = step()
barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    iq, ld = step()
e1.record()
barrier()
ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"config": f"BASELINE configs[4]: LowRankRootAddedDiag N={args.n} rank={args.rank} batch={args.batch} "
                                f"sharded over {world} GPU(s) ({nb} per GPU), fp32, cold calls",
                      "n_gpus": world, "ms_per_call": float(ms), "calls_per_s": 1e3 / float(ms),
                      "solves_per_s": args.batch * 1e3 / float(ms), "gathered": int(iq.numel()),
                      "inv_quad_mean": float(iq.mean()), "logdet_mean": float(ld.mean()),
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}))
if world > 1:
    dist.destroy_process_group()

"""Batch sharding across the GPUs of one box.

The Krylov path never couples batch elements except through the reference's batch-global stop tests, so the only
multi-GPU structure is: rank g owns a contiguous slice of the flattened batch, runs the whole path on it with no
data-path collective, and ONE all-gather of the per-rank results (inv_quad, logdet: two scalars per batch element)
assembles the answer.  N (the operator size) is never sharded.  One process per GPU, ``torch.distributed`` (NCCL over
NVLink on the GPU box, gloo in the CPU tests) is plumbing only.

Note on the stop rule: linear_cg stops on the mean residual over the WHOLE batch (utils/linear_cg.py:304 of the
reference).  With default settings every call runs exactly 21 iterations, so sharding cannot change the iteration
count; with a user tolerance each rank applies the rule to its own slice (documented deviation, +-1 iteration).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [start, stop) of a flattened batch for ``rank`` (first ``batch % world`` ranks get one
    extra element)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world: {rank}/{world}")
    base, extra = divmod(batch, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_results(inv_quad_local: torch.Tensor, logdet_local: torch.Tensor, batch: int,
                   group: Optional[dist.ProcessGroup] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """The single collective of the multi-GPU path: all-gather of the stacked per-rank (inv_quad, logdet) slices.
    Slices may be ragged (batch not divisible by world): they are padded to the largest slice and trimmed after."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return inv_quad_local, logdet_local
    rank = dist.get_rank(group)
    width = -(-batch // world)
    local = torch.zeros(2, width, dtype=inv_quad_local.dtype, device=inv_quad_local.device)
    n = inv_quad_local.numel()
    local[0, :n] = inv_quad_local.reshape(-1)
    local[1, :n] = logdet_local.reshape(-1)
    out = torch.empty(world * 2, width, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local, group=group)  # the one collective (ncclAllGather on the GPU box)
    out = out.view(world, 2, width)
    iq, ld = [], []
    for g in range(world):
        s, e = shard_bounds(batch, g, world)
        iq.append(out[g, 0, : e - s])
        ld.append(out[g, 1, : e - s])
    del rank
    return torch.cat(iq), torch.cat(ld)


def sharded_inv_quad_logdet(local_compute: Callable[[int, int], Tuple[torch.Tensor, torch.Tensor]], batch: int,
                            group: Optional[dist.ProcessGroup] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Runs ``local_compute(start, stop) -> (inv_quad, logdet)`` on this rank's slice of a flattened batch of size
    ``batch`` and returns the full-batch results on every rank.  ``local_compute`` builds the operator for its slice
    (inputs are generated / loaded per rank, there is no scatter) and calls ``op.inv_quad_logdet``."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    start, stop = shard_bounds(batch, rank, world)
    iq, ld = local_compute(start, stop)
    return gather_results(iq, ld, batch, group)


__all__ = ["shard_bounds", "gather_results", "sharded_inv_quad_logdet"]

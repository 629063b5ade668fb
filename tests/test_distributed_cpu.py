"""world_size-2 gloo test of the multi-GPU path on CPU: contiguous batch shards, no data-path collective, one
all-gather of the results.  The per-rank compute is the numpy oracle (the CUDA path needs a GPU); the sharding and the
collective are the product code under test."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from linear_operator_b200.distributed import gather_results, shard_bounds, sharded_inv_quad_logdet
from oracle import krylov_oracle as ko


def test_shard_bounds_cover_the_batch():
    for batch in (1, 7, 8, 1024, 4097):
        for world in (1, 2, 3, 8):
            pieces = [shard_bounds(batch, r, world) for r in range(world)]
            assert pieces[0][0] == 0 and pieces[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))
            sizes = [e - s for s, e in pieces]
            assert max(sizes) - min(sizes) <= 1


def _problem(batch=5, n=40, s=4):
    rng = np.random.default_rng(0)
    w = rng.standard_normal((batch, n, 8))
    A = w @ w.transpose(0, 2, 1) / 8
    d = np.full((batch, n), 0.5)
    rhs = rng.standard_normal((batch, n, 1))
    probes = rng.standard_normal((batch, n, s))
    probes /= np.linalg.norm(probes, axis=-2, keepdims=True)
    return A, d, rhs, probes


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    A, d, rhs, probes = _problem()

    def local(start, stop):
        iq, ld, _ = ko.dense_added_diag_inv_quad_logdet(A[start:stop], d[start:stop], rhs[start:stop],
                                                       probes=probes[start:stop])
        return torch.from_numpy(iq), torch.from_numpy(ld)

    iq, ld = sharded_inv_quad_logdet(local, A.shape[0])
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.stack([iq.numpy(), ld.numpy()]))
    dist.destroy_process_group()


def test_two_rank_sharded_inv_quad_logdet(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    A, d, rhs, probes = _problem()
    iq, ld, _ = ko.dense_added_diag_inv_quad_logdet(A, d, rhs, probes=probes)
    for r in range(2):  # every rank holds the full, ragged-batch (3 + 2) result
        got = np.load(tmp_path / f"r{r}.npy")
        # per-shard CG stops on the shard's own mean residual (documented); with default settings both run 21 iterations
        np.testing.assert_allclose(got[0], iq, rtol=1e-12)
        np.testing.assert_allclose(got[1], ld, rtol=1e-12)


def test_single_process_gather_is_identity():
    a, b = torch.arange(3.0), torch.arange(3.0) + 1
    x, y = gather_results(a, b, 3)
    assert x is a and y is b

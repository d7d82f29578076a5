"""world_size-2 gloo test of the N>1 host path: rank-range shards, allgather, sum allreduce, and the
Gray-slice partial sums of a single permanent (the oracle stands in for the per-rank device call)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lightworks_b200 import sharding
    from oracle import oracle as O  # noqa: N812

    U = O.random_unitary(7, 3)  # noqa: N806
    ins = [1, 1, 1, 1, 0, 0, 0]
    S = O.fock_count(7, 4)  # noqa: N806
    r0, r1 = sharding.rank_range(S, rank, world)
    local = torch.from_numpy(O.basis_probabilities(U, ins, r0, r1 - r0))  # per-rank shard (device call on a GPU box)
    full = sharding.allgather_shards(local, S, rank, world)
    total = sharding.allreduce_sum(local.sum().reshape(1).clone())
    ref = O.basis_probabilities(U, ins)
    ok = np.array_equal(full.numpy(), ref) and abs(float(total) - 1.0) < 1e-12
    dist.barrier()
    dist.destroy_process_group()
    ret[rank] = ok


@pytest.mark.timeout(120)
def test_rank_range_shards_allgather_gloo():
    world = 2
    port = 29500 + os.getpid() % 2000
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert all(ret[r] for r in range(world))

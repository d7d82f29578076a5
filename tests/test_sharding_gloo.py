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


def _piece_worker(rank, world, port, ret):
    """The multi-GPU pipeline of bench.py on gloo: equal-work shards in pieces of decreasing size, every rank writes
    piece j at the head of a max-length slot, one all-gather per piece, the distribution is the ragged view
    d_full[j].view(world, slot_j)[r, :len(piece j of rank r)]; plus the all-reduce of the sum."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lightworks_b200 import sharding
    from oracle import oracle as O  # noqa: N812

    m, n = 9, 7  # n >= 7: the work-balanced cuts are not the equal-count cuts
    U = O.random_unitary(m, 4)  # noqa: N806
    ins = [1] * n + [0] * (m - n)
    S = O.fock_count(m, n)  # noqa: N806
    pieces = sharding.work_balanced_ranges(m, n, world, [0.6, 0.3, 0.1])
    sub = len(pieces[0])
    slot_len = [max(pieces[r][j][1] - pieces[r][j][0] for r in range(world)) for j in range(sub)]
    d_slot = [torch.zeros(slot_len[j], dtype=torch.float64) for j in range(sub)]
    d_full = [torch.zeros(world * slot_len[j], dtype=torch.float64) for j in range(sub)]
    d_sum = torch.zeros(1, dtype=torch.float64)
    works = []
    for j in range(sub):
        a, b = pieces[rank][j]
        d_slot[j][: b - a] = torch.from_numpy(O.basis_probabilities(U, ins, a, b - a))  # the device call on a GPU box
        d_sum += d_slot[j].sum()
        works.append(dist.all_gather_into_tensor(d_full[j], d_slot[j], async_op=True))
    works.append(dist.all_reduce(d_sum, async_op=True))
    for w in works:
        w.wait()
    full = np.zeros(S)
    for j in range(sub):
        view = d_full[j].view(world, slot_len[j])
        for r in range(world):
            a, b = pieces[r][j]
            full[a:b] = view[r, : b - a].numpy()
            assert np.all(view[r, b - a:].numpy() == 0)  # padding stays zero
    ref = O.basis_probabilities(U, ins)
    counts = [pieces[r][-1][1] - pieces[r][0][0] for r in range(world)]
    ok = np.array_equal(full, ref) and abs(float(d_sum) - 1.0) < 1e-12 and len(set(counts)) > 1
    dist.barrier()
    dist.destroy_process_group()
    ret[rank] = ok


@pytest.mark.timeout(180)
def test_equal_work_pieces_pipeline_gloo():
    world = 2
    port = 31500 + os.getpid() % 2000
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_piece_worker, args=(world, port, ret), nprocs=world, join=True)
        assert all(ret[r] for r in range(world))


def _native_cuts_worker(rank, world, port, ret):
    """The product's multi-GPU step on gloo: shard boundaries from the C ABI (lwb_shard_cuts, what lwb_comm_* uses on
    a GPU box), every rank writes its shard into ITS slice of a contiguous S-long buffer, the exchange leaves every
    rank with the whole distribution in fock_basis order, and the all-reduced shard sums give the normalisation."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import lightworks_b200 as lb
    from oracle import oracle as O  # noqa: N812

    m, n = 9, 7
    U = O.random_unitary(m, 4)  # noqa: N806
    ins = [1] * n + [0] * (m - n)
    S = O.fock_count(m, n)  # noqa: N806
    cuts = lb.shard_cuts(m, n, [k / world for k in range(world + 1)])
    full = torch.zeros(S, dtype=torch.float64)
    a, b = cuts[rank], cuts[rank + 1]
    full[a:b] = torch.from_numpy(O.basis_probabilities(U, ins, a, b - a))  # the device kernels on a GPU box
    total = full[a:b].sum().reshape(1).clone()
    for r in range(world):  # in-place broadcasts of unequal length = the NCCL fallback of lwb_comm_perm_basis_probs_dev
        dist.broadcast(full[cuts[r]:cuts[r + 1]], src=r)
    dist.all_reduce(total)
    ref = O.basis_probabilities(U, ins)
    ok = (np.array_equal(full.numpy(), ref) and abs(float(total) - 1.0) < 1e-12 and cuts[0] == 0 and cuts[-1] == S
          and (cuts[1] - cuts[0]) != (cuts[2] - cuts[1]))  # equal work, not equal counts
    dist.barrier()
    dist.destroy_process_group()
    ret[rank] = ok


@pytest.mark.timeout(180)
def test_native_shard_cuts_contiguous_gather_gloo():
    world = 2
    port = 33500 + os.getpid() % 2000
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_native_cuts_worker, args=(world, port, ret), nprocs=world, join=True)
        assert all(ret[r] for r in range(world))

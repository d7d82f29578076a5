"""
Multi-GPU partitioning of the hot path (host logic; one process per GPU).

Output Fock states are independent units of identical cost, so the full distribution of one input is
sharded by contiguous RANK RANGES of the reference's fock_basis order; each rank unranks its own first
state on the device.  A single large permanent is sharded by Gray-code chunk ranges
(lwb_perm_single's slice/n_slices).  The only exchange steps are the allgather of probability shards /
the scalar allreduce of their sum, and a 16-byte allreduce of the partial permanents.
"""
from __future__ import annotations


def rank_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """[begin, end) of `total` units owned by `rank` of `world`; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank outside the world")
    return (total * rank) // world, (total * (rank + 1)) // world


def allgather_shards(local, total: int, rank: int, world: int, group=None):
    """All-gather rank-range shards (torch tensor, any backend) into one tensor of length `total`."""
    import torch
    import torch.distributed as dist

    sizes = [rank_range(total, r, world)[1] - rank_range(total, r, world)[0] for r in range(world)]
    pad = max(sizes)
    buf = torch.zeros(pad, dtype=local.dtype, device=local.device)
    buf[: local.numel()] = local
    out = [torch.empty(pad, dtype=local.dtype, device=local.device) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)])


def allreduce_sum(t, group=None):
    import torch.distributed as dist

    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t

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


def _average_terms(m: int, n: int) -> float:
    """Average terms of the multiplicity-aware Glynn walk (csrc/k1x.h) over all states of n photons in m modes."""
    from math import factorial

    def parts(k, mx):
        if k == 0:
            yield []
            return
        for p in range(min(k, mx), 0, -1):
            for rest in parts(k - p, p):
                yield [p, *rest]

    tot = work = 0
    for part in parts(n, n):
        d = len(part)
        if d > m:
            continue
        cnt = factorial(m) // factorial(m - d)
        for v in {x: part.count(x) for x in part}.values():
            cnt //= factorial(v)
        t = 1
        for s_ in part:
            t *= s_ + 1
        if any(s_ % 2 for s_ in part):
            t //= 2
        tot += cnt
        work += cnt * t
    return work / tot


def work_balanced_ranges(m: int, n: int, world: int) -> list[tuple[int, int]]:
    """Contiguous rank ranges of fock_basis(m, n) with (approximately) equal WORK for the multiplicity-aware
    permanent walk.  In the reference order the states whose highest occupied mode is <= v are the rank prefix
    [0, C(v+n, n)) and bunch more (fewer terms) than later ones; the cumulative work at those m breakpoints is
    exact (combinatorics over multiplicity patterns) and interpolated linearly in between."""
    from math import comb

    total = comb(m + n - 1, n)
    if world == 1 or n < 7:  # below n = 7 the expanded-row kernel is used: uniform cost per state
        return [rank_range(total, r, world) for r in range(world)]
    pts = [(0, 0.0)]
    for v in range(m):
        r = comb(v + n, n)
        pts.append((r, _average_terms(v + 1, n) * r))
    w_total = pts[-1][1]
    cuts = [0]
    for k in range(1, world):
        target = w_total * k / world
        for (r0, w0), (r1, w1) in zip(pts, pts[1:]):
            if w0 <= target <= w1:
                cuts.append(int(r0 + (r1 - r0) * (target - w0) / (w1 - w0)) if w1 > w0 else r0)
                break
    cuts.append(total)
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


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

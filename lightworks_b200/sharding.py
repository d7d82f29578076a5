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


def _colex_unrank(rank: int, n: int) -> list[int]:
    """Ascending mode list of the state at `rank` of fock_basis(m, n) (csrc/fock.cuh: rank = sum C(x_i + i - 1, i))."""
    from math import comb

    x = [0] * n
    for i in range(n, 0, -1):
        y = i - 1
        while comb(y + 1, i) <= rank:
            y += 1
        rank -= comb(y, i)
        x[i - 1] = y - i + 1
    return x


def _free_work(i: int, v: int) -> tuple[int, int]:
    """Over all ways to put i photons into v modes: (sum of prod(s_j + 1), the same sum restricted to states whose
    occupations are all even).  Generating functions (1 - z)^(-2v) and ((1 + z^2) / (1 - z^2)^2)^v."""
    from math import comb

    if v == 0:
        return (1, 1) if i == 0 else (0, 0)
    total = comb(i + 2 * v - 1, i)
    even = 0
    if i % 2 == 0:
        h = i // 2
        even = sum(comb(v, a) * comb(h - a + 2 * v - 1, 2 * v - 1) for a in range(min(v, h) + 1))
    return total, even


def cumulative_work(m: int, n: int, rank: int, overhead: float = 0.0) -> float:
    """Exact work of the multiplicity-aware walk (csrc/k1x.h) over the rank prefix [0, rank) of fock_basis(m, n):
    sum of terms(state) = prod(s_j + 1), halved when some occupation is odd, plus `overhead` term-equivalents per
    state (unranking and prologue).  The states below `rank` split, by the highest photon in which they differ from
    state(rank), into n blocks "photons above position i fixed, i photons free in the modes below y_i"; fixed and
    free photons occupy disjoint modes, so the block sums factorise (closed forms in _free_work)."""
    from math import comb

    total_states = comb(m + n - 1, n)
    if rank <= 0:
        return 0.0
    if rank >= total_states:
        a, e = _free_work(n, m)
        return (e + (a - e) // 2) + overhead * total_states
    y = _colex_unrank(rank, n)
    work = 0
    for i in range(n, 0, -1):  # states with x_j = y_j for j > i and x_i < y_i
        v = y[i - 1]           # free photons live in modes 0 .. v-1
        if v == 0:
            continue
        fixed = y[i:]
        f, odd = 1, False
        for mode in set(fixed):
            c = fixed.count(mode)
            f *= c + 1
            odd |= c % 2 == 1
        a, e = _free_work(i, v)
        work += f * a // 2 if odd else f * (e + (a - e) // 2)
    return work + overhead * rank


# per-state cost of classification, unranking and block prologue in units of one Glynn term: least-squares fit of
# time = a * terms + b * states to the 14 per-rank kernel times of the 2 / 4 / 8-GPU C3 runs on B200 (b / a = 152; with
# the round-1 constant 63 the rank holding the most, cheapest states ran 5 % longer than the others once the walk
# itself got faster)
STATE_OVERHEAD_TERMS = 150.0


def work_cuts(m: int, n: int, fractions: list[float]) -> list[int]:
    """Ranks of fock_basis(m, n) at which the cumulative work of the multiplicity-aware permanent walk reaches the
    given fractions of the total (ascending, in [0, 1]); bisection on the exact cumulative work function (host
    integer arithmetic, no device needed).  Outside 7 <= n <= 16 the expanded-row kernel runs: uniform cost per state.
    This is the readable statement of the arithmetic; the product path uses its C++ form lwb_shard_cuts
    (csrc/comm.cu), and tests/test_abi_and_host.py compares the two cut for cut."""
    from math import comb

    total = comb(m + n - 1, n)
    if n < 7 or n > 16 or total >= 2 ** 32:  # the multiplicity-aware walk covers 7 <= n <= 16 and < 2^32 states
        return [min(total, max(0, int(round(total * f)))) for f in fractions]
    w_total = cumulative_work(m, n, total, STATE_OVERHEAD_TERMS)
    cuts = []
    for f in fractions:
        if f <= 0:
            cuts.append(0)
            continue
        if f >= 1:
            cuts.append(total)
            continue
        target = w_total * f
        lo, hi = (cuts[-1] if cuts else 0), total
        while lo < hi:
            mid = (lo + hi) // 2
            if cumulative_work(m, n, mid, STATE_OVERHEAD_TERMS) < target:
                lo = mid + 1
            else:
                hi = mid
        cuts.append(lo)
    return cuts


def work_balanced_ranges(m: int, n: int, world: int, pieces: list[float] | None = None):
    """Contiguous rank ranges of fock_basis(m, n) with equal WORK per rank: bunched states (early ranks) need fewer
    terms, so equal-count shards would be up to ~1.3x out of balance.  With `pieces` (fractions of a shard summing
    to 1, e.g. [0.6, 0.3, 0.1]) every shard is returned as a list of sub-ranges holding those work fractions, for
    pipelines that exchange piece j while computing piece j+1."""
    if pieces is None:
        cuts = work_cuts(m, n, [k / world for k in range(world + 1)])
        return [(cuts[i], cuts[i + 1]) for i in range(world)]
    acc = [0.0]
    for p in pieces:
        acc.append(acc[-1] + p)
    fr = [(r + a / acc[-1]) / world for r in range(world) for a in acc[:-1]] + [1.0]
    cuts = work_cuts(m, n, fr)
    k = len(pieces)
    return [[(cuts[r * k + j], cuts[r * k + j + 1]) for j in range(k)] for r in range(world)]


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

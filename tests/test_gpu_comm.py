"""
Multi-GPU communicators behind the C ABI (include/lwb200.h "multi-GPU", csrc/comm.cu).

The single-GPU forms (LOCAL communicator over one device, RANK communicator of world size 1) run everywhere; the
tests that need two devices skip on a one-GPU box (the driver's GPU test box has one; they run under
`gpurun --gpus 2`).  Parity rule: a rank-range shard is bit-identical to the same slice of the stand-alone launch, so
every sharded result is compared with `==`, not with a tolerance.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle as O  # noqa: N812

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lb():
    import lightworks_b200

    return lightworks_b200


def _n_gpus():
    import torch

    return torch.cuda.device_count()


def _check_local(lb, n_gpus):
    ctx = lb.default_context(0)
    comm = lb.Comm.local(n_gpus)
    try:
        info = comm.info()
        assert info["world"] == n_gpus and info["kind"] == "local"
        for (m, n, seed) in ((9, 7, 4), (16, 8, 5), (12, 5, 2), (6, 9, 6)):
            U = O.random_unitary(m, seed)  # noqa: N806
            ins = np.bincount(np.random.default_rng(seed).integers(0, m, n), minlength=m).astype(np.uint8)
            ref = ctx.perm_basis_probs(U, ins)
            for peer in (1, 0):
                comm.set_option("peer_stores", peer)
                got, total = comm.perm_basis_probs(U, ins)
                assert np.array_equal(got, ref), (m, n, peer)
                assert abs(total - 1.0) < 1e-12
            comm.set_option("peer_stores", 1)
            cuts = [comm.shard(m, n, r) for r in range(n_gpus)]
            assert cuts[0][0] == 0 and cuts[-1][1] == len(ref) and all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            # full_probability_distribution: threshold, dict order, lossy marginalisation on the sharded probabilities
            keys, vals = comm.pdist(U, m, ins)
            rk, rv = ctx.pdist(U, m, ins, 1e-9, 0)
            assert np.array_equal(keys, rk) and np.array_equal(vals, rv)
            lossy = np.bincount(np.random.default_rng(seed + 1).integers(0, m - 3, n), minlength=m - 3).astype(np.uint8)
            keys, vals = comm.pdist(U, m - 3, lossy)  # 3 loss modes
            rk, rv = ctx.pdist(U, m - 3, lossy, 1e-9, 0)
            assert np.array_equal(keys, rk) and np.array_equal(vals, rv)
        A = O.random_unitary(60, 5)[:24, :24]  # noqa: N806
        want = sum(ctx.perm_single(A, slice_=s, n_slices=n_gpus) for s in range(n_gpus))
        assert comm.perm_single(A) == want            # n >= 24: Gray slices, partial sums added in device order
        hp = O.perm_hp(A)
        assert abs(comm.perm_single(A) - hp) <= 1e-10 * abs(hp)
        A = A[:21, :21]                               # below 24 one device walks the whole space (no exchange pays)
        assert comm.perm_single(A) == ctx.perm_single(A)
        # unique sub-inputs of an imperfect source, round-robin over the devices
        U = O.random_unitary(10, 4)  # noqa: N806
        subs = [[(b >> i) & 1 for i in range(5)] + [0] * 3 for b in range(32)]
        c0 = comm.context(0)
        many = c0.dists_from_circuit(U, 8, subs, 1e-9, 0, comm=comm)  # 2 loss modes
        for sb, d in zip(subs, many):
            one = c0.dist_from_circuit(U, 8, sb, 1e-9, 0)
            assert np.array_equal(d.to_numpy(), one.to_numpy()), sb
    finally:
        comm.close()


def test_local_communicator_one_gpu(lb):
    _check_local(lb, 1)


def test_local_communicator_all_gpus(lb):
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    _check_local(lb, min(_n_gpus(), 8))


def test_rank_communicator_world_1(lb):
    import torch

    ctx = lb.default_context(0)
    comm = lb.Comm.rank(ctx, 0, 1, lb.Comm.unique_id())
    try:
        U = O.random_unitary(12, 3)  # noqa: N806
        ins = [1] * 7 + [0] * 5
        ref = ctx.perm_basis_probs(U, ins)
        got, total = comm.perm_basis_probs(U, ins)
        assert np.array_equal(got, ref) and abs(total - 1.0) < 1e-12
        dU, dins = torch.from_numpy(U).cuda(), torch.tensor(ins, dtype=torch.uint8, device="cuda")  # noqa: N806
        for _ in range(3):  # alternates the two halves of the buffer
            p, s = comm.perm_basis_probs_dev(dU, 12, dins, 7)
            ctx.synchronize()
            assert np.array_equal(torch.as_tensor(p, device="cuda").cpu().numpy(), ref)
            assert abs(float(torch.as_tensor(s, device="cuda")[0]) - 1.0) < 1e-12
        A = O.random_unitary(60, 5)[:20, :20]  # noqa: N806
        assert comm.perm_single(A) == ctx.perm_single(A)
    finally:
        comm.close()


def test_drop_in_backend_with_n_gpus(lb):
    """set_n_gpus: the stand-alone backend shards full_probability_distribution over every visible GPU."""
    n = min(_n_gpus(), 8)
    U = O.random_unitary(12, 11)  # noqa: N806
    circ = lb.CompiledCircuitView(U, 10)  # 2 loss modes
    ins = [1, 1, 1, 1, 1, 1, 1, 0, 0, 0]
    ref = lb.PermanentBackend().full_probability_distribution(circ, ins)
    lb.set_n_gpus(n if n > 1 else 1)
    try:
        if n > 1:
            assert lb.backends.local_comm().info()["world"] == n
        got = lb.PermanentBackend().full_probability_distribution(circ, ins)
        assert list(got.keys()) == list(ref.keys()) and list(got.values()) == list(ref.values())
    finally:
        lb.set_n_gpus(1)


_RANK_SCRIPT = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
import lightworks_b200 as lb
from oracle import oracle as O
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
ctx = lb.Context(rank)
uid = [lb.Comm.unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
comm = lb.Comm.rank(ctx, rank, world, uid[0])
for peer in (1, 0):
    comm.set_option("peer_stores", peer)
    for (m, n, seed) in ((16, 8, 5), (9, 7, 4), (12, 5, 2), (20, 9, 8)):
        U = O.random_unitary(m, seed)
        ins = np.bincount(np.random.default_rng(seed).integers(0, m, n), minlength=m).astype(np.uint8)
        ref = ctx.perm_basis_probs(U, ins)
        got, total = comm.perm_basis_probs(U, ins)              # full copy on every rank's host
        assert np.array_equal(got, ref), (rank, m, n, peer)
        assert abs(total - 1.0) < 1e-12
        dU, dins = torch.from_numpy(U).cuda(), torch.from_numpy(ins).cuda()
        for rep in range(3):
            p, s = comm.perm_basis_probs_dev(dU, m, dins, n)
            ctx.synchronize()
            assert np.array_equal(torch.as_tensor(p, device="cuda").cpu().numpy(), ref), (rank, m, n, peer, rep)
        keys, vals = comm.pdist(U, m, ins)
        rk, rv = ctx.pdist(U, m, ins, 1e-9, 0)
        assert np.array_equal(keys, rk) and np.array_equal(vals, rv)
for n in (20, 24):
    A = O.random_unitary(60, 5)[:n, :n]
    hp = O.perm_hp(A)
    got = comm.perm_single(A)       # Gray slices + 16-byte all-reduce
    assert abs(got - hp) <= 1e-10 * abs(hp), (rank, n)
info = comm.info()
dist.barrier()
comm.close()
dist.destroy_process_group()
if rank == 0:
    print("RANK-COMM-OK", info)
"""


def test_rank_communicator_two_processes(lb, tmp_path):
    """One process per GPU, NCCL bootstrap through torch.distributed, CUDA IPC peer stores and the NCCL-broadcast
    exchange: every rank ends with the whole distribution, bit-identical to one GPU."""
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    script = tmp_path / "rank_comm.py"
    script.write_text(_RANK_SCRIPT.format(root=ROOT))
    world = min(_n_gpus(), 4)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "RANK-COMM-OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]

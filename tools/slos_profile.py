"""SLOS on C3 (24 modes, 10 photons): per-kernel timing target for ncu.  python tools/slos_profile.py [reps]"""
import sys

import torch

sys.path.insert(0, ".")
from lightworks_b200 import _lib as L  # noqa: E402
from oracle import oracle as O  # noqa: E402, N812

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
order = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = L.Context(0)
s = torch.cuda.Stream()
ctx.set_stream(s.cuda_stream)
with torch.cuda.stream(s):
    U = torch.from_numpy(O.random_unitary(24, 3)).cuda()
    S = L.fock_count(24, 10)
    p = torch.empty(S, dtype=torch.float64, device="cuda")
    ins = [1] * 10 + [0] * 14
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        ctx.slos_basis_probs_dev(U, 24, ins, p, order=order)
        e1.record(s)
        e1.synchronize()
        print(f"slos C3 order={order}: {e0.elapsed_time(e1):.3f} ms  sum={float(p.sum()):.12f}")

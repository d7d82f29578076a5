"""SLOS on C3 (24 modes, 10 photons): per-kernel timing target for ncu.
    python tools/slos_profile.py [reps] [order] [units 0|1]"""
import sys

import torch

sys.path.insert(0, ".")
from lightworks_b200 import _lib as L  # noqa: E402
from oracle import oracle as O  # noqa: E402, N812

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
order = int(sys.argv[2]) if len(sys.argv) > 2 else 1
units = int(sys.argv[3]) if len(sys.argv) > 3 else 1
L.set_option("slos_units", units)
ctx = L.Context(0)
s = torch.cuda.Stream()
ctx.set_stream(s.cuda_stream)
with torch.cuda.stream(s):
    import os
    shapes = ((24, 10, 3),) if os.environ.get("C3_ONLY") else ((24, 10, 3), (16, 8, 4), (12, 6, 2))
    for (m, n, seed) in shapes:
        U = torch.from_numpy(O.random_unitary(m, seed)).cuda()
        S = L.fock_count(m, n)
        p = torch.empty(S, dtype=torch.float64, device="cuda")
        ins = [1] * n + [0] * (m - n)
        best = 1e9
        for _ in range(reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            ctx.slos_basis_probs_dev(U, m, ins, p, order=order)
            e1.record(s)
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f"slos m={m} n={n} order={order} units={units}: {best:.3f} ms  sum={float(p.sum()):.12f}", flush=True)

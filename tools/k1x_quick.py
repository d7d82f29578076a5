"""Times the multiplicity-aware full-basis walk only (variant sweeps): C3, C4-sized and n=12."""
import sys

import torch

sys.path.insert(0, ".")
from lightworks_b200 import _lib as L  # noqa: E402
from oracle import oracle as O  # noqa: E402, N812

ctx = L.Context(0)
s = torch.cuda.Stream()
ctx.set_stream(s.cuda_stream)
with torch.cuda.stream(s):
    for (m, n, seed) in [(24, 10, 3), (16, 8, 4), (20, 12, 9), (18, 14, 7)]:
        U = torch.from_numpy(O.random_unitary(m, seed)).cuda()
        ins = torch.tensor([1] * n + [0] * (m - n), dtype=torch.uint8, device="cuda")
        S = min(L.fock_count(m, n), 100_000_000)
        p = torch.empty(S, dtype=torch.float64, device="cuda")
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            ctx.perm_basis_probs_dev(U, m, ins, n, 0, S, p)
            e1.record(s)
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f"{sys.argv[1]:>10s} m={m} n={n} S={S}: {best:.3f} ms  sum={float(p.sum()):.12f}")

"""C3 full distribution: multiplicity-aware walk vs expanded-row kernel."""
import sys

import torch

sys.path.insert(0, ".")
from lightworks_b200 import _lib as L  # noqa: E402
from oracle import oracle as O  # noqa: E402, N812

ctx = L.Context(0)
s = torch.cuda.Stream()
ctx.set_stream(s.cuda_stream)
with torch.cuda.stream(s):
    for (m, n, seed) in [(24, 10, 3), (16, 8, 4), (12, 6, 2), (20, 12, 9)]:
        U = torch.from_numpy(O.random_unitary(m, seed)).cuda()
        ins = torch.tensor([1] * n + [0] * (m - n), dtype=torch.uint8, device="cuda")
        S = min(L.fock_count(m, n), 100_000_000)
        p = torch.empty(S, dtype=torch.float64, device="cuda")
        q = torch.empty(S, dtype=torch.float64, device="cuda")
        for flag, buf in ((1, p), (0, q)):
            L.set_option("multiplicity_aware", flag)
            best = 1e9
            for rep in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(s)
                ctx.perm_basis_probs_dev(U, m, ins, n, 0, S, buf)
                e1.record(s)
                e1.synchronize()
                best = min(best, e0.elapsed_time(e1))
            print(f"m={m} n={n} S={S} multiplicity_aware={flag}: {best:.3f} ms  {S/best*1e3:.3e} probs/s  sum={float(buf.sum()):.12f}")
        err = ((p - q).abs() / q.clamp_min(1e-30))[q > 1e-14].max()
        print("   max rel diff:", float(err))
L.set_option("multiplicity_aware", 1)

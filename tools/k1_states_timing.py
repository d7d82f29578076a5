"""State-mode (basis rank range) K1 timing for several photon numbers and every lanes-per-permanent variant."""
import os
import sys

import torch

sys.path.insert(0, ".")
from lightworks_b200 import _lib as L  # noqa: E402
from oracle import oracle as O  # noqa: E402, N812

ctx = L.Context(0)
L.set_option("multiplicity_aware", 0)  # this tool times the expanded-row kernel variants
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
PEAK = float(os.environ.get("FP64_PEAK", "36.07e12"))
cases = [(16, 8), (20, 9), (24, 10), (24, 11), (26, 12), (28, 13), (30, 14), (36, 16), (60, 20)]
with torch.cuda.stream(stream):
    for m, n in cases:
        U = torch.from_numpy(O.random_unitary(m, n)).cuda()
        ins = torch.tensor([1] * n + [0] * (m - n), dtype=torch.uint8, device="cuda")
        S = L.fock_count(m, n)
        fl = 2.0 ** (n - 1) * (8 * n - 4)
        cnt = int(min(S, max(3e11 / fl, 148 * 64)))
        probs = torch.empty(cnt, dtype=torch.float64, device="cuda")
        for ll in range(0, 6):
            try:
                L.set_tuning(n, ll)
            except ValueError:
                continue
            best = 1e9
            for rep in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                ctx.perm_basis_probs_dev(U, m, ins, n, 0, cnt, probs)
                e1.record(stream)
                e1.synchronize()
                if rep:
                    best = min(best, e0.elapsed_time(e1) * 1e-3)
            print(f"states m={m} n={n} log_l={ll} cnt={cnt} t={best*1e3:.2f} ms {cnt/best:.3e} probs/s frac={cnt*fl/best/PEAK:.3f}", flush=True)
        L.set_tuning(n, -1)

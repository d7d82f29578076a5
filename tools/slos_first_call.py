"""First-call cost of the SLOS path at C3 (plan construction for the ten layers) vs a warm call."""
import sys
import time

import torch

sys.path.insert(0, ".")
from lightworks_b200 import _lib as L  # noqa: E402
from oracle import oracle as O  # noqa: E402, N812

ctx = L.Context(0)
m, n = 24, 10
U = torch.from_numpy(O.random_unitary(m, 3)).cuda()
p = torch.empty(L.fock_count(m, n), dtype=torch.float64, device="cuda")
ins = [1] * n + [0] * (m - n)
torch.cuda.synchronize()
for i in range(3):
    t0 = time.perf_counter()
    ctx.slos_basis_probs_dev(U, m, ins, p, order=1)
    ctx.synchronize()
    print(f"call {i}: {1e3 * (time.perf_counter() - t0):.2f} ms", flush=True)

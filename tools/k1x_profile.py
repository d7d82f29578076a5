import sys
import torch
sys.path.insert(0, ".")
from lightworks_b200 import _lib as L  # noqa: E402
from oracle import oracle as O  # noqa: E402, N812
ctx = L.Context(0)
s = torch.cuda.Stream()
ctx.set_stream(s.cuda_stream)
with torch.cuda.stream(s):
    U = torch.from_numpy(O.random_unitary(24, 3)).cuda()
    ins = torch.tensor([1] * 10 + [0] * 14, dtype=torch.uint8, device="cuda")
    S = L.fock_count(24, 10)
    p = torch.empty(S, dtype=torch.float64, device="cuda")
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        ctx.perm_basis_probs_dev(U, 24, ins, 10, 0, S, p)
        e1.record(s)
        e1.synchronize()
        print(f"{e0.elapsed_time(e1):.3f} ms sum={float(p.sum()):.12f}")

"""Single large permanents (K2) for ncu: python tools/k2_profile.py 24 26 28"""
import sys

import torch

sys.path.insert(0, ".")
from lightworks_b200 import _lib as L  # noqa: E402
from oracle import oracle as O  # noqa: E402, N812

ctx = L.Context(0)
s = torch.cuda.Stream()
ctx.set_stream(s.cuda_stream)
U60 = O.random_unitary(60, 5)
with torch.cuda.stream(s):
    for n in [int(v) for v in sys.argv[1:]] or [24]:
        A = torch.from_numpy(U60[:n, :n].copy()).cuda()
        o2 = torch.empty(2, dtype=torch.float64, device="cuda")
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            ctx.perm_single_dev(A, n, o2)
            e1.record(s)
            e1.synchronize()
        print(n, f"{e0.elapsed_time(e1) * 1e3:.1f} us", o2.tolist())

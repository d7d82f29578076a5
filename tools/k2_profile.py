"""Single large permanent (K2) timing / ncu target: python tools/k2_profile.py [n ...]"""
import sys

import torch

sys.path.insert(0, ".")
from lightworks_b200 import _lib as L  # noqa: E402
from oracle import oracle as O  # noqa: E402, N812

ns = [int(a) for a in sys.argv[1:]] or [24]
ctx = L.Context(0)
s = torch.cuda.Stream()
ctx.set_stream(s.cuda_stream)
U60 = O.random_unitary(60, 5)
with torch.cuda.stream(s):
    for n in ns:
        A = torch.from_numpy(U60[:n, :n].copy()).cuda()
        out = torch.zeros(2, dtype=torch.float64, device="cuda")
        best = 1e9
        for _ in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            ctx.perm_single_dev(A, n, out)
            e1.record(s)
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        # back-to-back launches: the per-launch share without the event pair
        reps = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(reps):
            ctx.perm_single_dev(A, n, out)
        e1.record(s)
        e1.synchronize()
        print(f"k2 n={n}: single launch {best * 1e3:.1f} us, back-to-back {e0.elapsed_time(e1) / reps * 1e3:.1f} us per launch, perm={complex(*out.cpu().tolist())}", flush=True)

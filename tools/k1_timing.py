"""Kernel timing sweep (device-resident inputs, CUDA events on the launch stream).
    LWB_LIB_PATH=... python tools/k1_timing.py [tag]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from lightworks_b200 import _lib as L  # noqa: E402
from oracle import oracle as O  # noqa: E402, N812

tag = sys.argv[1] if len(sys.argv) > 1 else "base"
ns = [int(x) for x in os.environ.get("NS", "8,10,12,16,20").split(",")]
ctx = L.Context(0)
L.set_option("multiplicity_aware", 0)  # this tool times the expanded-row kernel variants
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
PEAK = float(os.environ.get("FP64_PEAK", "36.07e12"))


def time_fn(fn, reps=4):
    with torch.cuda.stream(stream):
        fn()
        stream.synchronize()
        best = 1e9
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best


with torch.cuda.stream(stream):
    for n in ns:
        flops_per = 2.0 ** (n - 1) * (8 * n - 4)
        B = int(min(max(4e11 / flops_per, 148 * 64), 8e6))
        A = torch.randn(B, n, n, 2, dtype=torch.float64, device="cuda") / np.sqrt(n)
        out = torch.empty(B, 2, dtype=torch.float64, device="cuda")
        for ll in range(0, 6):
            try:
                L.set_tuning(n, ll)
                t = time_fn(lambda: ctx.perm_matrices_dev(A, B, n, out))
            except ValueError:
                continue
            tf = B * flops_per / t
            print(f"[{tag}] matrices n={n} log_l={ll} B={B} t={t*1e3:.3f} ms {B/t:.3e} perm/s {tf/1e12:.2f} TF frac={tf/PEAK:.3f}", flush=True)
        L.set_tuning(n, -1)
        del A, out

    if "C3" in os.environ.get("EXTRA", "C3"):
        U = torch.from_numpy(O.random_unitary(24, 3)).cuda()
        ins = torch.tensor([1] * 10 + [0] * 14, dtype=torch.uint8, device="cuda")
        S = L.fock_count(24, 10)
        cnt = S // 8
        probs = torch.empty(cnt, dtype=torch.float64, device="cuda")
        for ll in (1, 2, 3):
            L.set_tuning(10, ll)
            t = time_fn(lambda: ctx.perm_basis_probs_dev(U, 24, ins, 10, 0, cnt, probs), reps=3)
            tf = cnt * 2.0 ** 9 * 76 / t
            print(f"[{tag}] C3/8 log_l={ll} cnt={cnt} t={t*1e3:.2f} ms {cnt/t:.3e} probs/s {tf/1e12:.2f} TF frac={tf/PEAK:.3f}", flush=True)
        L.set_tuning(10, -1)
        U60 = O.random_unitary(60, 5)
        for n in [20, 22, 24, 26, 28, 30, 32]:
            A = torch.from_numpy(U60[:n, :n].copy()).cuda()
            o2 = torch.empty(2, dtype=torch.float64, device="cuda")
            t = time_fn(lambda: ctx.perm_single_dev(A, n, o2), reps=3)
            tf = 2.0 ** (n - 1) * (8 * n - 4) / t
            print(f"[{tag}] single n={n} t={t*1e3:.3f} ms {tf/1e12:.2f} TF frac={tf/PEAK:.3f}", flush=True)

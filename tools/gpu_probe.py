"""First-contact GPU probe: correctness spot checks against the oracle + kernel timings.
Run on the GPU box:  python tools/gpu_probe.py > gpurun_out/probe.log"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from lightworks_b200 import _lib as L  # noqa: E402
from oracle import oracle as O  # noqa: E402, N812

ctx = L.Context(0)
print(ctx.device_info())
peak = ctx.fp64_peak()
print("fp64 DFMA peak TFLOP/s:", peak / 1e12)

rng = np.random.default_rng(0)
# ---- correctness: matrices, every n, every variant ----
for n in range(1, 25):
    B = 67 if n <= 14 else 5
    A = (rng.normal(size=(B, n, n)) + 1j * rng.normal(size=(B, n, n))) / np.sqrt(n)
    ref = O.perm_batch(A, threads=8)
    for ll in range(0, 6):
        try:
            L.set_tuning(n, ll)
        except ValueError:
            continue
        got = ctx.perm_matrices(A)
        err = np.max(np.abs(got - ref) / np.abs(ref))
        got32 = ctx.perm_matrices(A, L.LWB_C64)
        err32 = np.max(np.abs(got32 - ref) / np.abs(ref))
        print(f"matrices n={n} log_l={ll} relerr c128={err:.2e} c64={err32:.2e}", "OK" if err < 1e-10 else "FAIL")
    L.set_tuning(n, -1)

# ---- correctness: basis probs ----
for (m, n, seed) in [(6, 3, 1), (12, 6, 2), (8, 5, 7), (10, 8, 3)]:
    U = O.random_unitary(m, seed)
    ins = [1] * n + [0] * (m - n)
    ref = O.basis_probabilities(U, ins, threads=8)
    got = ctx.perm_basis_probs(U, ins)
    print(f"basis m={m} n={n} maxabs={np.max(np.abs(got-ref)):.2e} sum={got.sum():.15f}",
          "OK" if np.allclose(got, ref, rtol=1e-10, atol=1e-16) else "FAIL")
    basis = ctx.fock_basis(m, n)
    print("  fock_basis bit-exact:", np.array_equal(basis, O.fock_basis(m, n)))
    amps = ctx.perm_amplitudes(U, [ins], basis)
    print("  amplitudes vs probs:", np.allclose(np.abs(amps[0]) ** 2, ref, rtol=1e-10, atol=1e-16))

# ---- correctness: single ----
U60 = O.random_unitary(60, 5)
for n in [2, 3, 5, 8, 12, 16, 20, 22]:
    A = U60[:n, :n]
    ref = O.perm(A)
    got = ctx.perm_single(A)
    parts = sum(ctx.perm_single(A, slice_=s, n_slices=3) for s in range(3))
    print(f"single n={n} relerr={abs(got-ref)/abs(ref):.2e} sliced={abs(parts-ref)/abs(ref):.2e}")

# ---- timing: matrices kernel (device resident) ----
def time_fn(fn, reps=5):
    torch.cuda.synchronize()
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best

ctx.set_stream(torch.cuda.current_stream().cuda_stream)
results = {}
for n in [6, 8, 10, 12, 14, 16, 18, 20]:
    flops_per = 2.0 ** (n - 1) * (8 * n - 4)
    B = int(min(max(2e10 / flops_per, 2000), 4e6))
    A = torch.randn(B, n, n, 2, dtype=torch.float64, device="cuda") / np.sqrt(n)
    out = torch.empty(B, 2, dtype=torch.float64, device="cuda")
    for ll in range(0, 6):
        try:
            L.set_tuning(n, ll)
        except ValueError:
            continue
        t = time_fn(lambda: ctx.perm_matrices_dev(A, B, n, out))
        tf = B * flops_per / t
        results[f"mat_n{n}_l{ll}"] = tf / peak
        print(f"matrices n={n} log_l={ll} B={B} t={t*1e3:.3f} ms  {B/t:.3e} perm/s  {tf/1e12:.2f} TFLOP/s  frac={tf/peak:.3f}")
    L.set_tuning(n, -1)

# ---- timing: C3 basis probs (m=24, n=10) on a slice ----
U = torch.from_numpy(O.random_unitary(24, 3)).cuda()
ins = torch.tensor([1] * 10 + [0] * 14, dtype=torch.uint8, device="cuda")
S = L.fock_count(24, 10)
cnt = S // 8
probs = torch.empty(cnt, dtype=torch.float64, device="cuda")
for ll in (3, 4, 5):
    L.set_tuning(10, ll)
    t = time_fn(lambda: ctx.perm_basis_probs_dev(U, 24, ins, 10, 0, cnt, probs), reps=3)
    tf = cnt * 2.0 ** 9 * 76 / t
    print(f"C3 slice log_l={ll} cnt={cnt} t={t*1e3:.2f} ms {cnt/t:.3e} probs/s {tf/1e12:.2f} TFLOP/s frac={tf/peak:.3f} sum={float(probs.sum()):.6f}")
L.set_tuning(10, -1)

# ---- timing: single ----
for n in [20, 24, 28, 30]:
    A = torch.from_numpy(U60[:n, :n].copy()).cuda()
    o2 = torch.empty(2, dtype=torch.float64, device="cuda")
    t = time_fn(lambda: ctx.perm_single_dev(A, n, o2), reps=3)
    tf = 2.0 ** (n - 1) * (8 * n - 4) / t
    print(f"single n={n} t={t*1e3:.3f} ms {tf/1e12:.2f} TFLOP/s frac={tf/peak:.3f}")
print(json.dumps(results))

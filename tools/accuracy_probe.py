"""Round-2 accuracy adjudication on the GPU box: which side of a GPU-vs-reference mismatch carries the rounding error?

For every size prints the relative error of (a) the CUDA kernels and (b) the reference algorithm restated in
oracle/lw_oracle.c (thewalrus perm_bbfg: one incremental double-precision loop) against the extended-precision
adjudicator oracle.perm_hp (long double, compensated).  Output goes to profiles/r02_accuracy_probe.log.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lightworks_b200 as lb  # noqa: E402
from oracle import oracle as O  # noqa: E402,N812

ctx = lb.default_context(0)
U60 = O.random_unitary(60, 5)
print("== single permanents, C5 matrices U60[:n,:n]: relative error against the long-double adjudicator ==")
print("n  |perm|  gpu_c128  reference_algorithm(double)  gpu_c64  gpu_sliced8  t_ref_s t_hp_s")
for n in (12, 16, 20, 22, 24, 26, 28):
    A = U60[:n, :n]
    t = time.time(); hp = O.perm_hp(A); t_hp = time.time() - t
    t = time.time(); ref = O.perm(A) if n <= 26 else None; t_ref = time.time() - t
    g = ctx.perm_single(A)
    g64 = ctx.perm_single(A, dtype=lb.LWB_C64)
    gs = sum(ctx.perm_single(A, slice_=s, n_slices=8) for s in range(8))
    e = lambda v: abs(v - hp) / abs(hp)  # noqa: E731
    print(n, f"{abs(hp):.3e} {e(g):.2e} {e(ref) if ref is not None else float('nan'):.2e} {e(g64):.2e} {e(gs):.2e} {t_ref:.2f} {t_hp:.2f}",
          flush=True)

print("== batched permanents (perm_matrices), 3 Haar blocks per n ==")
print("n  max gpu_c128 err  max reference_algorithm err  max gpu_c64 err")
for n in (8, 12, 13, 14, 15, 16, 17, 18, 19, 20):
    A = np.stack([O.random_unitary(n + 3, 100 * n + b)[:n, :n] for b in range(3)])
    hp = np.array([O.perm_hp(a) for a in A])
    ref = O.perm_batch(A, threads=3)
    g = ctx.perm_matrices(A)
    g64 = ctx.perm_matrices(A, dtype=lb.LWB_C64)
    e = lambda v: float(np.max(np.abs(v - hp) / np.abs(hp)))  # noqa: E731
    print(n, f"{e(g):.2e} {e(ref):.2e} {e(g64):.2e}", flush=True)

print("== full-basis probabilities: permanent kernels vs SLOS kernels (independent algorithms) ==")
for (m, n, seed, ins) in ((24, 10, 3, [1] * 10 + [0] * 14), (16, 8, 4, [1] * 8 + [0] * 8), (12, 6, 2, [1] * 6 + [0] * 6)):
    U = O.random_unitary(m, seed)
    p = ctx.perm_basis_probs(U, ins)
    s = ctx.slos_basis_probs(U, ins, order=0)
    p64 = ctx.perm_basis_probs(U, ins, dtype=lb.LWB_C64) if m < 24 else ctx.perm_basis_probs(U, ins, 10_000_000, 2_000_000, dtype=lb.LWB_C64)
    pref = p if m < 24 else p[10_000_000:12_000_000]
    big = p > 1e-7
    mid = (p > 1e-12) & ~big
    rel = np.abs(p - s) / np.maximum(p, 1e-300)
    big64 = pref > 1e-7
    print(f"m={m} n={n}: states {len(p)}, p>1e-7: {int(big.sum())} max rel {rel[big].max():.2e}; 1e-12<p<=1e-7: max rel "
          f"{rel[mid].max() if mid.any() else 0:.2e}; max abs {np.abs(p - s).max():.2e}; sum {p.sum():.15f} / {s.sum():.15f}; "
          f"c64 max rel (p>1e-7) {(np.abs(p64 - pref) / pref)[big64].max():.2e}", flush=True)
    cnt = min(len(p), 2000)
    ref = O.basis_probabilities(U, ins, 0, cnt, threads=8)
    k = ref > 1e-7
    print(f"   vs oracle (first {cnt}): max rel perm {np.max(np.abs(p[:cnt] - ref)[k] / ref[k]) if k.any() else 0:.2e}  slos "
          f"{np.max(np.abs(s[:cnt] - ref)[k] / ref[k]) if k.any() else 0:.2e}", flush=True)

"""complex64 error statistics of the full-basis kernels (float multiplicity-aware walk and float expanded-row kernel)
against the complex128 result, for bunched and collision-free inputs: max |dp| / p over p > 1e-7 and max |dp| / max(p, mean p)."""
import sys

import numpy as np

sys.path.insert(0, ".")
import lightworks_b200 as lb  # noqa: E402
from oracle import oracle as O  # noqa: E402,N812

ctx = lb.default_context(0)
print("m n input kernel : max rel (p>1e-7) | max |dp|/max(p,mean) | max |dp|/max(p)")
for (m, n, seed) in [(9, 7, 4), (16, 8, 5), (6, 9, 6), (14, 10, 7), (5, 12, 8), (4, 16, 9), (20, 7, 10), (24, 10, 3), (20, 12, 9)]:
    U = O.random_unitary(m, seed)
    rng = np.random.default_rng(seed)
    inputs = {"bunched": np.bincount(rng.integers(0, m, n), minlength=m).astype(np.uint8)}
    if m >= n:
        inputs["unbunched"] = np.array([1] * n + [0] * (m - n), dtype=np.uint8)
    S = lb.fock_count(m, n)
    r0, cnt = S // 7, min(S - S // 7, 400000)
    for name, ins in inputs.items():
        ref = ctx.perm_basis_probs(U, ins, r0, cnt)
        for aware in (1, 0):
            lb.set_option("multiplicity_aware", aware)
            got = ctx.perm_basis_probs(U, ins, r0, cnt, dtype=1)
            lb.set_option("multiplicity_aware", 1)
            d = np.abs(got - ref)
            big = ref > 1e-7
            print(m, n, name, "walk " if aware else "plain", f": {(d[big] / ref[big]).max() if big.any() else 0:.2e} | "
                  f"{(d / np.maximum(ref, ref.mean())).max():.2e} | {d.max() / ref.max():.2e}", flush=True)

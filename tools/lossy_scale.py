"""C4' scale check: lossy 16-mode circuit (16 loss modes), 8 photons -> 61.5 M full states marginalised to 735 k keys."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from lightworks_b200 import _lib as L  # noqa: E402
from oracle import oracle as O  # noqa: E402, N812

ctx = L.Context(0)
rng = np.random.default_rng(0)
for (nm, n) in [(6, 4), (16, 6), (16, 8)]:
    m_tot = 2 * nm
    U = O.random_unitary(m_tot, nm)  # any unitary on n_modes + loss modes exercises the same path
    ins = [1] * n + [0] * (nm - n)
    for backend in (0, 1):
        t0 = time.perf_counter()
        keys, vals = ctx.pdist(U, nm, ins, 1e-9, backend)
        dt = time.perf_counter() - t0
        print(f"n_modes={nm} n={n} backend={backend}: {len(vals)} keys, sum={vals.sum():.12f}, {dt*1e3:.1f} ms "
              f"(S_full={L.fock_count(m_tot, n)})", flush=True)
        if nm == 6:
            rk, rv = (O.pdist_permanent if backend == 0 else O.pdist_slos)(U, nm, ins)
            print("   vs oracle: keys equal", np.array_equal(keys, rk), "max rel", np.max(np.abs(vals - rv) / rv))

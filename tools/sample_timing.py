"""Timing breakdown of on-device sampling on C3 (host API)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import lightworks_b200 as lb  # noqa: E402
from oracle import oracle as O  # noqa: E402, N812

ctx = lb.default_context(0)
m, n = 24, 10
U = O.random_unitary(m, 3)
ins = [1] * n + [0] * (m - n)
for ns in (10, 1000, 1_000_000, 1_000_000, 10_000_000):
    u = np.random.default_rng(0).random(ns)
    t0 = time.perf_counter()
    idx, tot = ctx.basis_sample(U, ins, u)
    t1 = time.perf_counter()
    print(f"basis_sample n_samples={ns}: {1e3 * (t1 - t0):.2f} ms total={tot:.12f}")
for be in (0, 1):
    t0 = time.perf_counter()
    idx, tot = ctx.basis_sample(U, ins, u[:1000], backend=be)
    print(f"backend {be}: {1e3 * (time.perf_counter() - t0):.2f} ms")

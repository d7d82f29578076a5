"""Small-configuration pass over every kernel family for compute-sanitizer: C1 (6 modes,
3 photons), a reduced C2 (10 modes, 5 photons), a 7-photon case for the multiplicity-aware walk, lossy distributions,
SLOS, sampling, the distribution algebra and a 14 x 14 single permanent.  Every result is also checked against the
oracle, so a sanitizer-clean run is a correct run.
    compute-sanitizer --tool memcheck  python tools/sanitize_small.py
    compute-sanitizer --tool racecheck python tools/sanitize_small.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lightworks_b200 as lb  # noqa: E402
from oracle import oracle as O  # noqa: E402,N812

ctx = lb.default_context(0)
ok = lambda a, b, tol=1e-10: np.allclose(a, b, rtol=tol, atol=1e-17)  # noqa: E731
U6, U10, U8 = O.random_unitary(6, 1), O.random_unitary(10, 2), O.random_unitary(8, 7)
assert ok(ctx.perm_basis_probs(U6, [1, 1, 1, 0, 0, 0]), O.basis_probabilities(U6, [1, 1, 1, 0, 0, 0]))
ins10 = [1] * 5 + [0] * 5
assert ok(ctx.perm_basis_probs(U10, ins10), O.basis_probabilities(U10, ins10))
assert ok(ctx.perm_basis_probs(U10, ins10, dtype=lb.LWB_C64), O.basis_probabilities(U10, ins10), 1e-4)
ins8 = [2, 1, 0, 1, 1, 0, 1, 1]
assert ok(ctx.perm_basis_probs(U8, ins8), O.basis_probabilities(U8, ins8, threads=4))
assert np.array_equal(ctx.fock_basis(6, 3), O.fock_basis(6, 3))
for backend, fn in ((0, O.pdist_permanent), (1, O.pdist_slos)):
    k, v = ctx.pdist(U10, 10, ins10, 1e-9, backend)
    rk, rv = fn(U10, 10, ins10)
    assert np.array_equal(k, rk) and ok(v, rv)
    k, v = ctx.pdist(U10, 7, [1, 1, 1, 0, 1, 0, 0], 1e-9, backend)  # 3 loss modes
    rk, rv = fn(U10, 7, [1, 1, 1, 0, 1, 0, 0])
    assert np.array_equal(k, rk) and np.allclose(v, rv, rtol=1e-9, atol=1e-15)
A = np.stack([O.random_unitary(12, s)[:9, :9] for s in range(40)])
assert ok(ctx.perm_matrices(A), O.perm_batch(A))
outs = O.fock_basis(6, 3)[::3]
assert ok(ctx.perm_amplitudes(U6, [[1, 1, 1, 0, 0, 0], [0, 2, 0, 1, 0, 0]], outs),
          O.amplitude_matrix(U6, [[1, 1, 1, 0, 0, 0], [0, 2, 0, 1, 0, 0]], outs))
A14 = O.random_unitary(20, 5)[:14, :14]
assert abs(ctx.perm_single(A14) - O.perm(A14)) <= 1e-10 * abs(O.perm(A14))
assert abs(sum(ctx.perm_single(A14, slice_=s, n_slices=3) for s in range(3)) - O.perm(A14)) <= 1e-10 * abs(O.perm(A14))
u = np.random.default_rng(3).random(500)
idx, _ = ctx.basis_sample(U8, ins8, u)
ref8 = O.basis_probabilities(U8, ins8, threads=4)
assert np.count_nonzero(idx != O.sample_indices(np.where(ref8 > 1e-9, ref8, 0.0), u)) <= 1
circ = lb.CompiledCircuitView(U6, 6)
inputs = {((0,), (0,), (1,), (), (), ()): 0.7, ((0,), (1,), (2,), (), (), ()): 0.3}
got = lb.pdist_calc(circ, inputs, lb.PermanentBackend())
ref = O.py_annotated_pdist_calc(circ, inputs, O.pdist_func("permanent"))
gm = {tuple(int(v) for v in k): float(p) for k, p in zip(got.keys_array, got.vals_array)}
assert set(gm) == set(ref) and all(abs(gm[k] - v) <= 1e-10 * v for k, v in ref.items())
print("sanitize_small: all checks passed, launches =", ctx.launch_count())

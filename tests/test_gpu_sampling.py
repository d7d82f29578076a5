"""
GPU parity tests of on-device sampling: lwb_sample / lwb_basis_sample through
the C ABI against the sampling oracle (numpy's Generator.choice, the reference's own sampler dependency,
simulation/sampler.py:329-349).  Index work: bit-exact, except that a uniform variate within rounding distance
of a CDF edge may land on the neighbouring non-empty state (the parallel scan sums in another order than
numpy's sequential cumsum); the tests bound that to 1e-5 of the samples and check the neighbour property.
"""
from collections import Counter

import numpy as np
import pytest

from oracle import oracle as O  # noqa: N812

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lb():
    import lightworks_b200

    return lightworks_b200


@pytest.fixture(scope="module")
def ctx(lb):
    return lb.default_context(0)


def _check(p, u, got, threshold=-1.0):
    ref = O.sample_indices(p, u, threshold)
    bad = np.nonzero(ref != got)[0]
    assert len(bad) <= max(0, int(1e-5 * len(u))), f"{len(bad)} of {len(u)} samples differ"
    kept = np.nonzero(np.asarray(p) > threshold)[0]
    for k in bad:  # an edge case may only move to the neighbouring state that has probability
        a, b = np.searchsorted(kept, [ref[k], got[k]])
        assert abs(int(a) - int(b)) == 1
    assert np.all(np.asarray(p)[got.astype(np.int64)] > threshold)  # never a dropped / zero-probability state


@pytest.mark.parametrize("S,n,seed", [(1, 100, 0), (2, 1000, 1), (5, 1000, 2), (4096, 20000, 3), (4097, 20000, 4),
                                      (100003, 100000, 5), ((1 << 22) + 7, 200000, 6)])
def test_sample_matches_numpy_choice(ctx, S, n, seed):  # noqa: N803
    rng = np.random.default_rng(1000 + seed)
    p = rng.random(S) ** 4
    p[rng.random(S) < 0.3] = 0.0  # zero-probability states, also at the ends
    if S > 4:
        p[0] = p[-1] = 0.0
    if not p.any():
        p[0] = 1.0
    p /= p.sum()
    u = np.random.default_rng(seed).random(n)
    got, total = ctx.sample(p, u)
    assert total == pytest.approx(1.0, rel=1e-12)
    _check(p, u, got)


def test_sample_unnormalised_and_threshold(ctx):
    rng = np.random.default_rng(7)
    p = rng.random(50000) * 3.7  # numpy normalises by the total; so does the device CDF
    p[::3] = 1e-12
    u = np.random.default_rng(8).random(30000)
    got, total = ctx.sample(p, u, threshold=1e-9)
    assert total == pytest.approx(p[p > 1e-9].sum(), rel=1e-12)
    _check(p, u, got, threshold=1e-9)


def test_sample_edges(ctx):
    # no samples requested; a single state; extreme variates
    idx, total = ctx.sample([0.25, 0.75], np.zeros(0))
    assert len(idx) == 0 and total == 1.0
    idx, _ = ctx.sample([0.0, 2.0, 0.0], [0.0, 0.5, np.nextafter(1.0, 0.0)])
    assert idx.tolist() == [1, 1, 1]
    idx, _ = ctx.sample([0.5, 0.5], [0.0, np.nextafter(0.5, 0.0), 0.5, np.nextafter(1.0, 0.0)])
    assert idx.tolist() == [0, 0, 1, 1]  # side='right': u == cdf edge belongs to the next state
    with pytest.raises(ValueError):  # nothing above the threshold (numpy: "probabilities contain NaN")
        ctx.sample([1e-12, 1e-12], [0.3], threshold=1e-9)
    with pytest.raises(ValueError):
        ctx.sample(np.zeros(0), [0.3])


@pytest.mark.parametrize("backend", ["permanent", "slos"])
@pytest.mark.parametrize("m,n,seed", [(6, 3, 1), (12, 6, 2), (10, 7, 5)])
def test_basis_sample_is_the_reference_sampler(lb, backend, m, n, seed):
    """full_probability_distribution + _sample_N_outputs + _gen_samples_from_dist (sampler.py:240-349) on the
    oracle vs the same chain run on the device; dict order = the backend's own state order."""
    U = O.random_unitary(m, seed)  # noqa: N806
    ins = [1] * n + [0] * (m - n)
    b = lb.PermanentBackend() if backend == "permanent" else lb.SLOSBackend()
    circuit = lb.CompiledCircuitView(U, m)
    ref_dist = (O.pdist_permanent if backend == "permanent" else O.pdist_slos)(U, m, ins)
    keys, vals = ref_dist if isinstance(ref_dist, tuple) else (None, None)
    if keys is not None:
        ref_dist = {tuple(int(v) for v in k): float(p) for k, p in zip(keys, vals)}
    n_samples = 20000
    ref = O.sample_states(ref_dist, n_samples, seed=seed + 40)
    got = b.sample_outputs(circuit, ins, n_samples, seed=seed + 40)
    got = {tuple(k.s if hasattr(k, "s") else k): c for k, c in got.items()}
    ref = {tuple(k.s if hasattr(k, "s") else k): c for k, c in ref.items()}
    assert sum(got.values()) == n_samples
    common = sum(min(got.get(k, 0), ref.get(k, 0)) for k in set(got) | set(ref))
    assert common >= n_samples - 2  # identical up to a CDF-edge variate
    assert list(got)[:50] == list(ref)[:50] or common < n_samples  # same first-drawn order when identical


def test_c3_sampler_on_device(lb, ctx):
    """BASELINE config 3 as a Sampler: 24 modes, 10 photons, 92 561 040 outputs; the distribution never leaves the
    GPU.  Properties: drawn ranks are valid, reproducible for a seed, and their empirical probabilities agree with
    the probabilities of the drawn states themselves."""
    m, n = 24, 10
    U = O.random_unitary(m, 3)  # noqa: N806
    ins = [1] * n + [0] * (m - n)
    u = np.random.default_rng(11).random(200000)
    idx, total = ctx.basis_sample(U, ins, u, threshold=1e-9)
    idx2, _ = ctx.basis_sample(U, ins, u, threshold=1e-9)
    S = lb._lib.fock_count(m, n)  # noqa: N806
    assert np.array_equal(idx, idx2) and idx.max() < S
    assert 0.9 < total <= 1.0 + 1e-12  # states below the threshold are dropped, as in the reference's dict
    # the most frequently drawn ranks must be high-probability states: check them through the pair API
    top = [r for r, _ in Counter(idx.tolist()).most_common(20)]
    outs = np.array([lb._lib.fock_unrank(m, n, r) for r in top], dtype=np.uint8)
    p = ctx.perm_amplitudes(U, np.array([ins], dtype=np.uint8), outs, probs=True)[0]
    assert np.all(p > 1e-9)  # never a state the threshold drops
    # monotone variates draw monotone ranks (the CDF search is order preserving)
    us = np.sort(u[:5000])
    idx3, _ = ctx.basis_sample(U, ins, us, threshold=1e-9)
    assert np.all(np.diff(idx3.astype(np.int64)) >= 0)

"""
GPU parity tests of the dense-distribution algebra: lwb_dist_from_circuit,
lwb_dist_convolve, lwb_dist_axpy and lightworks_b200.convolution.pdist_calc against
  * the distributions the unmodified reference's pdist_calc returned (tests/golden/ref_convolve.json),
  * the oracle restatement (oracle.py_pdist_calc / py_annotated_pdist_calc) on seeded inputs,
  * a brute-force numpy convolution of random dense vectors.
Tolerance: rel 1e-10 (complex128 path) on every entry, abs 1e-16; the SET of states must be identical (the order
of the returned entries is the key-space order by design, see lightworks_b200/convolution.py).
"""
import itertools
import json
import os

import numpy as np
import pytest

from oracle import oracle as O  # noqa: N812

pytestmark = pytest.mark.gpu
REL = 1e-10


@pytest.fixture(scope="module")
def lb():
    import lightworks_b200

    return lightworks_b200


@pytest.fixture(scope="module")
def ctx(lb):
    return lb.default_context(0)


def _records():
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_convolve.json")) as f:
        return json.load(f)["records"]


class _Circ:
    def __init__(self, U, n_modes):  # noqa: N803
        self.U_full = np.asarray(U, dtype=np.complex128)
        self.n_modes = n_modes
        self.loss_modes = self.U_full.shape[0] - n_modes


def _as_dict(dist):
    return {tuple(int(v) for v in k): float(p) for k, p in zip(dist.keys_array, dist.vals_array)}


def _same(got: dict, ref: dict, rel=REL, abs_=1e-16, vacuum_abs=None):
    assert set(got) == set(ref), (sorted(set(got) ^ set(ref))[:5], len(got), len(ref))
    for k, v in ref.items():
        tol = rel * abs(v) + abs_
        if vacuum_abs is not None and sum(k) == 0:
            tol = vacuum_abs
        assert abs(got[k] - v) <= tol, (k, got[k], v)


@pytest.mark.parametrize("ka,kb,m", [(1, 1, 3), (2, 3, 4), (0, 2, 3), (3, 0, 5), (2, 2, 6)])
def test_convolve_matches_bruteforce(lb, ctx, ka, kb, m):
    L = lb._lib  # noqa: N806
    rng = np.random.default_rng(ka * 10 + kb)
    A = rng.random(L.keyspace_size(m, ka))  # noqa: N806
    B = rng.random(L.keyspace_size(m, kb))  # noqa: N806
    ref = np.zeros(L.keyspace_size(m, ka + kb))
    sa = L.keyspace_states(m, ka, np.arange(len(A)))
    sb = L.keyspace_states(m, kb, np.arange(len(B)))
    for i, j in itertools.product(range(len(A)), range(len(B))):
        ref[L.keyspace_index(sa[i] + sb[j])] += A[i] * B[j]
    da, db = ctx.dist_from_dense(m, ka, A), ctx.dist_from_dense(m, kb, B)
    out = (da @ db).to_numpy()
    assert np.allclose(out, ref, rtol=1e-13, atol=0)
    # commutative, and the accumulator is a prefix axpy
    assert np.allclose((db @ da).to_numpy(), ref, rtol=1e-13, atol=0)
    acc = ctx.dist_zeros(m, ka + kb + 1)
    acc.add_scaled(0.25, da @ db)
    acc.add_scaled(2.0, da)
    exp = np.zeros(L.keyspace_size(m, ka + kb + 1))
    exp[: len(ref)] += 0.25 * ref
    exp[: len(A)] += 2.0 * A
    assert np.allclose(acc.to_numpy(), exp, rtol=1e-13, atol=0)
    with pytest.raises(ValueError):
        ctx.dist_zeros(m, 0).add_scaled(1.0, da) if ka > 0 else (_ for _ in ()).throw(ValueError())
    with pytest.raises(ValueError):
        _ = da @ ctx.dist_zeros(m + 1, 1)


def test_axpy_many_equals_sequential_axpys(lb, ctx):
    """lwb_dist_axpy_many: one launch, the same operations per element in the same order as one lwb_dist_axpy per term
    (bit-identical), with terms over different prefix key spaces; argument errors."""
    L, m = lb._lib, 5  # noqa: N806
    rng = np.random.default_rng(8)
    kmaxs = [0, 3, 1, 2, 3, 0, 2]
    xs = [ctx.dist_from_dense(m, k, rng.random(L.keyspace_size(m, k))) for k in kmaxs]
    alphas = rng.random(len(xs)) * 3 - 1
    start = rng.random(L.keyspace_size(m, 4))
    one, many = ctx.dist_from_dense(m, 4, start), ctx.dist_from_dense(m, 4, start)
    for a, x in zip(alphas, xs):
        one.add_scaled(float(a), x)
    many.add_scaled_many(alphas, xs)
    assert np.array_equal(one.to_numpy(), many.to_numpy())
    many.add_scaled_many([], [])  # nothing to do
    assert np.array_equal(one.to_numpy(), many.to_numpy())
    with pytest.raises(ValueError):
        ctx.dist_zeros(m, 2).add_scaled_many([1.0], [xs[1]])  # term larger than the accumulator
    with pytest.raises(ValueError):
        many.add_scaled_many([1.0], [many])  # the accumulator cannot be its own term
    with pytest.raises(ValueError):
        many.add_scaled_many([1.0, 2.0], [xs[0]])


@pytest.mark.parametrize("backend", ["permanent", "slos"])
@pytest.mark.parametrize("case", ["lossless", "lossy", "vacuum"])
def test_dist_from_circuit_is_full_probability_distribution(lb, ctx, backend, case):
    L = lb._lib  # noqa: N806
    bid = L.BACKEND_PERMANENT if backend == "permanent" else L.BACKEND_SLOS
    if case == "lossy":
        r = next(r for r in _records() if r["loss_modes"] > 0 and r["annotated"])
        U, n_modes, ins = np.array(r["U_re"]) + 1j * np.array(r["U_im"]), r["n_modes"], [1, 0, 1, 0]  # noqa: N806
    else:
        U, n_modes = O.random_unitary(7, 21), 7  # noqa: N806
        ins = [2, 0, 1, 0, 1, 0, 0] if case == "lossless" else [0] * 7
    d = ctx.dist_from_circuit(U, n_modes, ins, 1e-9, bid)
    dense = d.to_numpy()
    keys, vals = (O.pdist_permanent if backend == "permanent" else O.pdist_slos)(U, n_modes, ins, 1e-9)
    ref = {tuple(int(v) for v in k): float(p) for k, p in zip(keys, vals)}
    idx = np.nonzero(dense)[0]
    got = {tuple(int(v) for v in k): float(dense[i]) for k, i in zip(L.keyspace_states(n_modes, d.kmax, idx), idx)}
    _same(got, ref, vacuum_abs=1e-12)


@pytest.mark.parametrize("r", _records(), ids=lambda r: f"{r['scenario'][:40]}[{r['backend']}]")
def test_pdist_calc_matches_reference(lb, r):
    """The reference's own pdist_calc outputs for its Source-generated inputs (both CPU backends)."""
    from lightworks_b200.convolution import pdist_calc

    U = np.array(r["U_re"]) + 1j * np.array(r["U_im"])  # noqa: N806
    circ = _Circ(U, r["n_modes"])
    b = lb.PermanentBackend() if r["backend"] == "permanent" else lb.SLOSBackend()
    if r["annotated"]:
        inputs = {tuple(tuple(mode) for mode in s): p for s, p in r["inputs"]}
    else:
        inputs = {tuple(s): p for s, p in r["inputs"]}
    got = _as_dict(pdist_calc(circ, inputs, b))
    ref = {tuple(k): v for k, v in zip(r["keys"], r["vals"])}
    # vacuum of lossy circuits: 1 - sum (permanent.py:159-161, probability_distribution.py:70-73), compared absolutely
    _same(got, ref, vacuum_abs=1e-12)
    assert sum(got.values()) == pytest.approx(sum(ref.values()), abs=1e-12)


def test_pdist_calc_matches_oracle_partial_distinguishability(lb):
    """8 modes, 4 photons: every set partition of the photons into distinguishable groups (the structure
    Source._build_statistics produces for indistinguishability < 1), weights from a seeded RNG."""
    from lightworks_b200.convolution import pdist_calc

    m, photons = 8, [0, 2, 4, 7]
    U = O.random_unitary(m, 31)  # noqa: N806
    circ = _Circ(U, m)

    def partitions(items):
        if not items:
            yield []
            return
        head, rest = items[0], items[1:]
        for p in partitions(rest):
            for i in range(len(p)):
                yield p[:i] + [[head] + p[i]] + p[i + 1:]
            yield [[head]] + p

    rng = np.random.default_rng(5)
    inputs = {}
    for part in partitions(photons):
        modes = [[] for _ in range(m)]
        for label, group in enumerate(part):
            for ph in group:
                modes[ph].append(label)
        inputs[tuple(tuple(x) for x in modes)] = float(rng.random())
    tot = sum(inputs.values())
    inputs = {k: v / tot for k, v in inputs.items()}
    assert len(inputs) == 15  # Bell(4)
    ref = O.py_annotated_pdist_calc(circ, inputs, O.pdist_func("permanent"))
    res = pdist_calc(circ, inputs, lb.PermanentBackend())
    _same(_as_dict(res), ref)
    assert sum(ref.values()) == pytest.approx(1.0, abs=1e-9)
    assert res.n_convolutions < sum(len(set(lab for md in k for lab in md)) - 1 for k in inputs)  # shared prefixes


def test_device_distribution_sampling(lb, ctx):
    U = O.random_unitary(6, 3)  # noqa: N806
    d = ctx.dist_from_circuit(U, 6, [1, 1, 0, 1, 0, 0]) @ ctx.dist_from_circuit(U, 6, [0, 0, 1, 0, 0, 0])
    dense = d.to_numpy()
    u = np.random.default_rng(2).random(20000)
    idx, total = d.sample(u)
    assert total == pytest.approx(dense.sum(), rel=1e-12)
    ref = O.sample_indices(dense, u)
    assert np.count_nonzero(ref != idx) <= 1


def test_sample_source_outputs_statistics(lb):
    """Whole noisy Sampler chain on the device (source statistics -> convolution -> sampling -> detector model): the
    empirical mean photon number per mode against the exact expectation from the oracle's distribution of the same
    imperfect source (efficiency-thinned), within 5 standard errors; and reproducibility for a seed."""
    m, ins = 6, [1, 0, 1, 1, 0, 0]
    U = O.random_unitary(m, 17)  # noqa: N806
    circ = _Circ(U, m)
    kw = dict(purity=0.95, brightness=0.85, indistinguishability=0.9)
    _, inputs = O.py_build_statistics(ins, **kw)
    ref = O.py_annotated_pdist_calc(circ, inputs, O.pdist_func("permanent"))
    eff, n = 0.8, 40000
    expect = eff * sum(p * np.array(k, dtype=float) for k, p in ref.items())
    second = sum(p * (eff * (1 - eff) * np.array(k, dtype=float) + (eff * np.array(k, dtype=float)) ** 2) for k, p in ref.items())
    sigma = np.sqrt(np.maximum(second - expect**2, 1e-12) / n)
    b = lb.PermanentBackend()
    got = lb.sample_source_outputs(circ, ins, n, b, detector_efficiency=eff, seed=5, **kw)
    assert sum(got.values()) == n
    mean = sum(c * np.array(k.s, dtype=float) for k, c in got.items()) / n
    assert np.all(np.abs(mean - expect) < 5 * sigma + 1e-9), (mean, expect, sigma)
    again = lb.sample_source_outputs(circ, ins, n, b, detector_efficiency=eff, seed=5, **kw)
    assert {tuple(k.s): c for k, c in got.items()} == {tuple(k.s): c for k, c in again.items()}
    clicks = lb.sample_source_outputs(circ, ins, 2000, b, photon_counting=False, p_dark=0.01, seed=6, **kw)
    assert all(max(k.s) <= 1 for k in clicks)

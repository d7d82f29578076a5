"""
CPU tests of the HOST orchestration in lightworks_b200/convolution.py (grouping of annotated inputs by their largest
photon group, canonical product chains and their cache, the vacuum rule) with the device replaced by a numpy stand-in
that implements the same dense key-space algebra by brute force on top of the oracle.  The stand-in is test
infrastructure: the product path has no CPU fallback (tests/test_abi_and_host.py checks that it raises without a GPU).
"""
import itertools
import json
import os

import numpy as np
import pytest

import lightworks_b200 as lb
from lightworks_b200 import _lib
from lightworks_b200 import convolution as conv
from oracle import oracle as O  # noqa: N812


class _FakeDist:
    def __init__(self, n_modes, kmax, dense):
        self.n_modes, self.kmax, self.dense = n_modes, kmax, np.asarray(dense, dtype=float)
        self.size = len(self.dense)

    def __matmul__(self, other):
        kmax = self.kmax + other.kmax
        out = np.zeros(_lib.keyspace_size(self.n_modes, kmax))
        sa = _lib.keyspace_states(self.n_modes, self.kmax, np.nonzero(self.dense)[0])
        sb = _lib.keyspace_states(self.n_modes, other.kmax, np.nonzero(other.dense)[0])
        pa, pb = self.dense[np.nonzero(self.dense)[0]], other.dense[np.nonzero(other.dense)[0]]
        for (i, a), (j, b) in itertools.product(enumerate(sa), enumerate(sb)):
            out[_lib.keyspace_index(a + b)] += pa[i] * pb[j]
        return _FakeDist(self.n_modes, kmax, out)

    def add_scaled(self, alpha, x):
        assert x.n_modes == self.n_modes and x.kmax <= self.kmax
        self.dense[: x.size] += alpha * x.dense

    def add_scaled_many(self, alphas, xs):
        for alpha, x in zip(alphas, xs):
            self.add_scaled(alpha, x)

    def to_numpy(self):
        return self.dense.copy()


class _FakeCtx:
    def dist_zeros(self, n_modes, kmax):
        return _FakeDist(n_modes, kmax, np.zeros(_lib.keyspace_size(n_modes, kmax)))

    def dist_from_dense(self, n_modes, kmax, dense):
        return _FakeDist(n_modes, kmax, dense)

    def dists_from_circuit(self, U_full, n_modes, in_states, threshold, backend, dtype, comm=None):  # noqa: N803
        return [self.dist_from_circuit(U_full, n_modes, s, threshold, backend, dtype) for s in in_states]

    def dist_from_circuit(self, U_full, n_modes, in_state, threshold, backend, dtype):  # noqa: N803
        fn = O.pdist_permanent if backend == _lib.BACKEND_PERMANENT else O.pdist_slos
        keys, vals = fn(U_full, n_modes, list(in_state), threshold)
        kmax = int(sum(in_state))
        dense = np.zeros(_lib.keyspace_size(n_modes, kmax))
        for k, p in zip(keys, vals):
            dense[_lib.keyspace_index(k)] = p
        return _FakeDist(n_modes, kmax, dense)


class _FakeBackend:
    _dtype = _lib.LWB_C128

    def __init__(self, name):
        self._backend_id = _lib.BACKEND_PERMANENT if name == "permanent" else _lib.BACKEND_SLOS
        self.ctx = _FakeCtx()


class _Circ:
    def __init__(self, r):
        self.U_full = np.array(r["U_re"]) + 1j * np.array(r["U_im"])
        self.n_modes, self.loss_modes = r["n_modes"], r["loss_modes"]


def _records():
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_convolve.json")) as f:
        return [r for r in json.load(f)["records"] if len(r["keys"]) <= 100]  # brute-force stand-in: small cases


@pytest.mark.parametrize("r", _records(), ids=lambda r: f"{r['scenario'][:40]}[{r['backend']}]")
def test_pdist_calc_orchestration_matches_reference(r):
    if r["annotated"]:
        inputs = {tuple(tuple(mode) for mode in s): p for s, p in r["inputs"]}
    else:
        inputs = {tuple(s): p for s, p in r["inputs"]}
    got = conv.pdist_calc(_Circ(r), inputs, _FakeBackend(r["backend"]))
    got = {tuple(int(v) for v in k): float(p) for k, p in zip(got.keys_array, got.vals_array)}
    ref = {tuple(k): v for k, v in zip(r["keys"], r["vals"])}
    assert set(got) == set(ref)
    for k, v in ref.items():
        assert got[k] == pytest.approx(v, rel=1e-10, abs=1e-12 if sum(k) == 0 else 1e-16)


def test_labels_to_states_and_grouping():
    # probability_distribution.py:106-121: one occupation vector per label, first-seen label order
    assert conv._labels_to_states(((0,), (0, 1), (), (2,)), 4) == [(1, 1, 0, 0), (0, 1, 0, 0), (0, 0, 0, 1)]
    assert conv._labels_to_states(((), ()), 2) == [(0, 0)]
    assert conv._is_annotated(((0,), ())) and not conv._is_annotated((1, 0))
    # chains are canonical (commutative product): the cache sees permutations of the same groups as one product
    alg = conv.DistributionAlgebra(_FakeCtx(), lb.CompiledCircuitView(O.random_unitary(4, 1)), _lib.BACKEND_PERMANENT, 1e-9)
    a, b, c = (1, 0, 0, 0), (0, 1, 0, 0), (0, 0, 2, 0)
    d1 = alg.product([c, a, b])
    n1 = alg.n_convolutions
    d2 = alg.product([b, c, a])
    assert d2 is d1 and alg.n_convolutions == n1 == 2
    alg.product([a, b])  # a prefix of the cached chain
    assert alg.n_convolutions == 2


def test_work_balanced_pieces_partition():
    from lightworks_b200 import sharding

    m, n, world = 24, 10, 4
    total = lb.fock_count(m, n)
    pieces = sharding.work_balanced_ranges(m, n, world, [0.6, 0.3, 0.1])
    flat = [p for shard in pieces for p in shard]
    assert flat[0][0] == 0 and flat[-1][1] == total and all(a[1] == b[0] for a, b in zip(flat, flat[1:]))
    c0 = sharding.STATE_OVERHEAD_TERMS
    w = [[sharding.cumulative_work(m, n, b, c0) - sharding.cumulative_work(m, n, a, c0) for a, b in shard] for shard in pieces]
    per_rank = [sum(x) for x in w]
    assert max(per_rank) / (sum(per_rank) / world) < 1.0001
    for x in w:
        assert x[0] / sum(x) == pytest.approx(0.6, abs=1e-4) and x[2] / sum(x) == pytest.approx(0.1, abs=1e-4)
    assert sharding.work_balanced_ranges(6, 3, 2, [0.5, 0.5]) == [[(0, 14), (14, 28)], [(28, 42), (42, 56)]]  # uniform below n = 7


def test_source_classes_path_equals_label_path():
    """pdist_calc fed with SourceClasses (arrays from lwb_source_statistics) against the same statistics as a dict of
    annotated label tuples: identical distributions (host orchestration on the numpy stand-in)."""
    m, ins = 5, [1, 1, 0, 2, 0]
    circ = lb.CompiledCircuitView(O.random_unitary(m, 23))
    kw = dict(purity=0.9, brightness=0.8, indistinguishability=0.85, probability_threshold=1e-5)
    for backend in ("permanent", "slos"):
        ann, entries = _lib.source_statistics(ins, **kw)
        assert ann
        a = conv.pdist_calc(circ, dict(entries), _FakeBackend(backend))
        b = conv.pdist_calc(circ, conv.SourceClasses.from_source(ins, **kw), _FakeBackend(backend))
        assert np.array_equal(a.keys_array, b.keys_array)
        assert np.allclose(a.vals_array, b.vals_array, rtol=1e-12, atol=0)
        # and against the oracle's restatement of the whole reference chain
        _, inputs = O.py_build_statistics(ins, **kw)
        ref = O.py_annotated_pdist_calc(circ, inputs, O.pdist_func(backend))
        got = {tuple(int(v) for v in k): float(p) for k, p in zip(b.keys_array, b.vals_array)}
        assert set(got) == set(ref) and all(got[k] == pytest.approx(v, rel=1e-10) for k, v in ref.items())
    # brightness-only source: plain State inputs through the same class
    basic = conv.SourceClasses.from_source([1, 0, 1, 0, 0], brightness=0.7)
    assert not basic.annotated and len(basic) == 4
    d = conv.pdist_calc(lb.CompiledCircuitView(O.random_unitary(5, 3)), basic, _FakeBackend("permanent"))
    assert d.vals_array.sum() == pytest.approx(1.0, abs=1e-8)


def test_detector_and_counting_helpers():
    """detect(): thinning / dark counts / threshold detection of whole sample arrays against their exact moments;
    count_rows_first_seen(): Counter semantics (first-seen order) on the packed-key and the byte-string path."""
    from collections import Counter

    from lightworks_b200.convolution import count_rows_first_seen, detect

    rng = np.random.default_rng(3)
    base = np.array([[2, 0, 1, 0, 3], [0, 0, 0, 1, 1], [1, 1, 1, 1, 1]], dtype=np.uint8)
    rows = base[rng.integers(0, 3, 60000)]
    eff = 0.7
    out = detect(rows.copy(), np.random.default_rng(4), eff)
    assert out.dtype == np.uint8 and out.shape == rows.shape and np.all(out <= rows)
    mean, expect = out.mean(axis=0), eff * rows.mean(axis=0)
    var = (eff * (1 - eff) * rows + (eff * rows.astype(float)) ** 2).mean(axis=0) - expect**2
    assert np.all(np.abs(mean - expect) < 5 * np.sqrt(var / len(rows)) + 1e-9)
    assert np.array_equal(detect(rows.copy(), np.random.default_rng(4), 1.0), rows)          # perfect detector: untouched
    dark = detect(np.zeros((40000, 4), dtype=np.uint8), np.random.default_rng(5), 1.0, p_dark=0.05)
    assert abs(dark.mean() - 0.05) < 5 * np.sqrt(0.05 * 0.95 / dark.size)
    clicks = detect(rows.copy(), np.random.default_rng(6), 0.9, p_dark=0.01, photon_counting=False)
    assert clicks.max() <= 1
    for data in (out, np.concatenate([out, out], axis=1).astype(np.uint8) * 9):  # 5 x 2 bits packed; 10 x 5 bits: byte strings
        first, counts = count_rows_first_seen(data)
        ref = Counter(map(tuple, data.tolist()))
        assert [tuple(r) for r in data[first].tolist()] == list(ref.keys())
        assert counts.tolist() == list(ref.values())
    f0, c0 = count_rows_first_seen(np.zeros((0, 3), dtype=np.uint8))
    assert len(f0) == 0 and len(c0) == 0

"""Property tests (hypothesis) of the exact host-side integer work: Fock / SLOS / key-space indexing round trips and
order contracts (emulator/utils/state.py:22-35, backends/slos.py:96-103), and the cumulative-work function."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

import lightworks_b200 as lb
from lightworks_b200 import _lib, sharding

mn = st.tuples(st.integers(1, 24), st.integers(0, 12))


@settings(max_examples=200, deadline=None)
@given(mn, st.data())
def test_fock_rank_unrank_roundtrip_and_order(m_n, data):
    m, n = m_n
    total = lb.fock_count(m, n)
    r = data.draw(st.integers(0, total - 1))
    s = lb.fock_unrank(m, n, r)
    assert int(s.sum()) == n and lb.fock_rank(s) == r
    if r + 1 < total:  # reference order: ascending on (s[m-1], ..., s[1]) (state.py:27-35)
        t = lb.fock_unrank(m, n, r + 1)
        assert tuple(s[::-1][:-1]) < tuple(t[::-1][:-1]) or m == 1


@settings(max_examples=200, deadline=None)
@given(mn, st.data())
def test_slos_rank_unrank_roundtrip_and_order(m_n, data):
    m, n = m_n
    total = lb.fock_count(m, n)
    r = data.draw(st.integers(0, total - 1))
    s = lb.slos_unrank(m, n, r)
    assert int(s.sum()) == n and lb.slos_rank(s) == r
    if r + 1 < total:  # dict insertion order of slos.py:96-103 = descending lexicographic occupation vectors
        t = lb.slos_unrank(m, n, r + 1)
        assert tuple(s) > tuple(t)


@settings(max_examples=200, deadline=None)
@given(st.integers(1, 16), st.integers(0, 8), st.data())
def test_keyspace_roundtrip(m, kmax, data):
    size = _lib.keyspace_size(m, kmax)
    i = data.draw(st.integers(0, size - 1))
    s = _lib.keyspace_state(m, kmax, i)
    assert _lib.keyspace_index(s) == i and int(s.sum()) <= kmax
    k = int(s.sum())  # blocks by photon number, fock_basis order inside a block
    assert i == _lib.keyspace_size(m, k - 1) + lb.fock_rank(s) if k > 0 else i == 0


@settings(max_examples=100, deadline=None)
@given(st.tuples(st.integers(2, 20), st.integers(7, 12)), st.data())
def test_cumulative_work_monotone_and_bounded(m_n, data):
    m, n = m_n
    total = lb.fock_count(m, n)
    a = data.draw(st.integers(0, total))
    b = data.draw(st.integers(a, total))
    wa, wb = sharding.cumulative_work(m, n, a), sharding.cumulative_work(m, n, b)
    assert 0 <= wa <= wb <= sharding.cumulative_work(m, n, total)
    assert wb - wa >= (b - a)  # at least one term per state
    assert wb - wa <= (b - a) * 2 ** (n - 1) * (1 if n > 1 else 2)  # at most the expanded-row term count


def test_work_cuts_match_first_state_terms():
    # the work of the single state at rank r is cumulative_work(r + 1) - cumulative_work(r)
    m, n = 9, 7
    rng = np.random.default_rng(0)
    for r in rng.integers(0, lb.fock_count(m, n), size=50):
        s = lb.fock_unrank(m, n, int(r))
        occ = [int(v) for v in s if v]
        t = int(np.prod([v + 1 for v in occ]))
        if any(v % 2 for v in occ):
            t //= 2
        assert sharding.cumulative_work(m, n, int(r) + 1) - sharding.cumulative_work(m, n, int(r)) == t

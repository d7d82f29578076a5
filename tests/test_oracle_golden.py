"""
Pins the CPU oracle (oracle/lw_oracle.c + oracle/oracle.py) against
  (a) the literals the reference's own tests hold for the hot path, and
  (b) outputs of the unmodified reference recorded in tests/golden/ (oracle/make_golden.py).
CPU only.
"""
import numpy as np
import pytest

from oracle import oracle as O  # noqa: N812
from tests.golden_util import golden, rel_close

G = golden()


@pytest.mark.parametrize("lit", [l for l in G.literals if l.get("kind") in ("amp", "prob", "slos")],
                         ids=lambda l: l["cite"])
def test_reference_literals(lit):
    U = G.arr(lit["U"])  # noqa: N806
    exp = complex(*lit["expected"])
    if lit["kind"] == "amp":
        got_c = O.probability_amplitude(U, lit["in"], lit["out"])
        got_py = O.py_probability_amplitude(U, lit["in"], lit["out"])
    elif lit["kind"] == "prob":
        got_c = abs(O.probability_amplitude(U, lit["in"], lit["out"])) ** 2
        got_py = abs(O.py_probability_amplitude(U, lit["in"], lit["out"])) ** 2
    else:
        keys, amps = O.slos_calculate(U, lit["in"])
        got_c = amps[[tuple(k) for k in keys.tolist()].index(tuple(lit["out"]))]
        got_py = O.py_slos_calculate(U, lit["in"])[tuple(lit["out"])]
    assert abs(got_c - exp) <= lit["rel"] * abs(exp)
    assert abs(got_py - exp) <= lit["rel"] * abs(exp)


def test_slos_hom_exact_zero():
    # tests/emulator/backend_test.py:236  r[1,1] == 0 exactly
    lit = next(l for l in G.literals if l.get("kind") == "slos_exact_zero")
    U = G.arr(lit["U"])  # noqa: N806
    keys, amps = O.slos_calculate(U, lit["in"])
    assert amps[[tuple(k) for k in keys.tolist()].index((1, 1))] == 0
    assert O.py_slos_calculate(U, lit["in"])[(1, 1)] == 0


def test_pdist_literal_backend_test_61_68():
    lit = next(l for l in G.literals if "pdist" in l)
    U = O.random_unitary(3, 2)  # noqa: N806
    for fn in (O.pdist_permanent, O.pdist_slos):
        keys, vals = fn(U, 3, [1, 1, 0])
        assert len(keys) == 6
        for k, v in zip(keys.tolist(), vals):
            assert abs(lit["pdist"][str(k)] - v) <= 1e-8 * v


def test_random_unitary_pinned():
    # tests/sdk/util_test.py:53-69 pins random_unitary(4, seed=111); first row is enough to pin scipy's stream
    U = O.random_unitary(4, 111)  # noqa: N806
    assert np.allclose(U @ U.conj().T, np.eye(4), atol=1e-12)
    rec = next(r for r in G.of_kind("amp") if "backend_test.py:153" in r["scenario"])
    assert np.array_equal(G.arr(rec["U"]), O.random_unitary(6, 23))


@pytest.mark.parametrize("rec", G.of_kind("fock"), ids=lambda r: f"m{r['m']}n{r['n']}")
def test_fock_basis_bit_exact(rec):
    ref = G.arr(rec["basis"])
    assert np.array_equal(O.fock_basis(rec["m"], rec["n"]), ref)
    assert O.fock_count(rec["m"], rec["n"]) == len(ref)
    if len(ref) <= 2000:
        assert O.py_fock_basis(rec["m"], rec["n"]) == ref.tolist()


@pytest.mark.parametrize("idx", range(len(G.of_kind("amp"))))
def test_amplitudes_match_reference(idx):
    rec = G.of_kind("amp")[idx]
    U, ins, outs, amps = (G.arr(rec[k]) for k in ("U", "ins", "outs", "amps"))  # noqa: N806
    got = np.array([O.probability_amplitude(U, i, o) for i, o in zip(ins, outs)])
    # identical algorithm in C and numba/numpy: agreement is at rounding level
    assert rel_close(got, amps, 1e-12, 1e-15), rec["scenario"]
    for k in range(min(5, len(ins))):
        assert abs(O.py_probability_amplitude(U, ins[k].tolist(), outs[k].tolist()) - amps[k]) <= 1e-12 * abs(amps[k]) + 1e-15


@pytest.mark.parametrize("idx", range(len(G.of_kind("pdist"))))
def test_pdist_matches_reference(idx):
    rec = G.of_kind("pdist")[idx]
    U = G.arr(rec["U"])  # noqa: N806
    fn = O.pdist_permanent if rec["backend"] == "permanent" else O.pdist_slos
    keys, vals = fn(U, rec["n_modes"], rec["in"], rec["threshold"])
    assert np.array_equal(keys, G.arr(rec["keys"])), rec["scenario"]  # same states, same dict order
    assert rel_close(vals, G.arr(rec["vals"]), 1e-11, 1e-15), rec["scenario"]
    if len(keys) <= 300:
        pyfn = O.py_pdist_permanent if rec["backend"] == "permanent" else O.py_pdist_slos
        d = pyfn(U, rec["n_modes"], rec["in"], rec["threshold"])
        assert [list(k) for k in d] == keys.tolist()
        assert rel_close(list(d.values()), G.arr(rec["vals"]), 1e-11, 1e-15)


@pytest.mark.parametrize("idx", range(len(G.of_kind("slos"))))
def test_slos_calculate_matches_reference(idx):
    rec = G.of_kind("slos")[idx]
    U = G.arr(rec["U"])  # noqa: N806
    keys, amps = O.slos_calculate(U, rec["in"])
    assert np.array_equal(keys, G.arr(rec["keys"]))
    assert rel_close(amps, G.arr(rec["amps"]), 1e-12, 1e-16)


def test_simulator_records_match_oracle():
    for rec in G.of_kind("simulator"):
        # task-level array = amplitudes of (input+heralds+loss zeros) -> (output+...) ; no heralds in these scenarios
        arr = G.arr(rec["array"])
        ins, outs = G.arr(rec["inputs"]), G.arr(rec["outputs"])
        amp_rec = [r for r in G.of_kind("amp") if r["scenario"] == rec["scenario"]]
        assert amp_rec, rec["scenario"]
        U = G.arr(amp_rec[0]["U"])  # noqa: N806
        pad = U.shape[0] - ins.shape[1]
        fi = np.pad(ins, ((0, 0), (0, pad)))
        fo = np.pad(outs, ((0, 0), (0, pad)))
        assert rel_close(O.amplitude_matrix(U, fi, fo), arr, 1e-12, 1e-15)


def test_glynn_branch_vs_definition_and_slos():
    # n>=4 is not pinned by a reference literal: cross-check the restated Glynn loop
    # against the permutation-sum definition and against the independent in-tree SLOS algorithm.
    rng = np.random.default_rng(5)
    for n in range(4, 9):
        A = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))  # noqa: N806
        assert abs(O.perm(A) - O.py_perm_naive(A)) <= 1e-12 * abs(O.py_perm_naive(A))
    U = O.random_unitary(7, 3)  # noqa: N806
    ins = [1, 1, 2, 0, 1, 0, 1]
    keys, amps = O.slos_calculate(U, ins)
    for k in range(0, len(keys), 37):
        assert abs(O.probability_amplitude(U, ins, keys[k]) - amps[k]) <= 1e-12


def test_mt_drivers_equal_serial():
    U = O.random_unitary(8, 1)  # noqa: N806
    a = O.basis_probabilities(U, [1, 1, 1, 1, 0, 0, 0, 0], 10, 200)
    b = O.basis_probabilities(U, [1, 1, 1, 1, 0, 0, 0, 0], 10, 200, threads=3)
    assert np.array_equal(a, b)
    rng = np.random.default_rng(0)
    A = rng.normal(size=(17, 6, 6)) + 1j * rng.normal(size=(17, 6, 6))  # noqa: N806
    assert np.array_equal(O.perm_batch(A), O.perm_batch(A, threads=4))


@pytest.mark.parametrize("S,n,seed", [(1, 10, 0), (7, 1000, 1), (4097, 20000, 2), (100003, 50000, 3)])
def test_sample_indices_is_numpy_choice(S, n, seed):  # noqa: N803
    """The sampling oracle against the reference's own sampler dependency: rng.choice (sampler.py:346-347)."""
    rng = np.random.default_rng(100 + seed)
    p = rng.random(S) ** 4
    p[rng.random(S) < 0.2] = 0.0
    if not p.any():
        p[0] = 1.0
    p /= p.sum()
    ref = np.random.default_rng(seed).choice(range(S), p=p, size=n)
    got = O.sample_indices(p, np.random.default_rng(seed).random(n))
    assert np.array_equal(ref, got.astype(ref.dtype))
    # the dict form the reference uses, with normalize=True
    dist = {i: float(v) for i, v in enumerate(p) if v > 0}
    counts = O.sample_states(dist, n, seed)
    keys = np.array(list(dist))
    ref2 = np.random.default_rng(seed).choice(range(len(dist)), p=np.array(list(dist.values())) / sum(dist.values()), size=n)
    assert counts == {int(k): int(c) for k, c in zip(*np.unique(keys[ref2], return_counts=True))}


def _convolve_records():
    import json
    import os

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_convolve.json")) as f:
        return json.load(f)["records"]


class _Circ:
    def __init__(self, r):
        self.U_full = np.array(r["U_re"]) + 1j * np.array(r["U_im"])
        self.n_modes, self.loss_modes = r["n_modes"], r["loss_modes"]


@pytest.mark.parametrize("r", _convolve_records(), ids=lambda r: f"{r['scenario'][:40]}[{r['backend']}]")
def test_pdist_calc_restatement_matches_reference(r):
    """oracle.py_pdist_calc / py_annotated_pdist_calc (probability_distribution.py:25-164) on the inputs the
    reference's Source produced, against the distribution the unmodified reference returned: same states in the
    same dict order, probabilities to 1e-12."""
    circ = _Circ(r)
    if r["annotated"]:
        inputs = {tuple(tuple(mode) for mode in s): p for s, p in r["inputs"]}
        got = O.py_annotated_pdist_calc(circ, inputs, O.pdist_func(r["backend"]))
    else:
        inputs = {tuple(s): p for s, p in r["inputs"]}
        got = O.py_pdist_calc(circ, inputs, O.pdist_func(r["backend"]))
    assert [list(k) for k in got] == r["keys"]
    assert rel_close(list(got.values()), r["vals"], 1e-12, 1e-18)


def _source_records():
    import json
    import os

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_source.json")) as f:
        return json.load(f)["records"]


@pytest.mark.parametrize("r", _source_records(), ids=lambda r: f"{r['state']}-{sorted(r['source'].items())}"[:60])
def test_source_statistics_match_reference(r):
    """Source._build_statistics (components/source.py:134-161) of the unmodified reference against (1) the oracle
    restatement and (2) the native enumeration lwb_source_statistics: same inputs in the same dict order, same
    State / AnnotatedState kind, probabilities to 1e-12."""
    import lightworks_b200._lib as L  # noqa: N812

    kw = {"purity": 1.0, "brightness": 1.0, "indistinguishability": 1.0, "probability_threshold": 1e-9, **r["source"]}
    want_keys = [tuple(tuple(m) for m in k) if r["annotated"] else tuple(k) for k in r["keys"]]
    ann, got = O.py_build_statistics(r["state"], **kw)
    assert ann == r["annotated"] and list(got) == want_keys
    assert rel_close(list(got.values()), r["vals"], 1e-12)
    for string_keys in (0, 1):  # packed integer keys, and the string-key path used when they would not fit 62 bits
        L.set_option("source_string_keys", string_keys)
        try:
            ann2, entries = L.source_statistics(r["state"], **kw)
        finally:
            L.set_option("source_string_keys", 0)
        assert ann2 == r["annotated"] and [k for k, _ in entries] == want_keys
        assert rel_close([p for _, p in entries], r["vals"], 1e-12)


def test_source_statistics_errors_and_scale():
    import time

    import lightworks_b200._lib as L  # noqa: N812

    with pytest.raises(ValueError):  # components/source.py:442-443
        L.source_statistics([1, 0], purity=0.4)
    with pytest.raises(ValueError):
        L.source_statistics([1, 0], brightness=1.5)
    # C4: 16 modes, 8 photons, Source(0.99, 0.9, 0.95): 26 068 annotated inputs; the reference needs
    # 131 s of Python for this enumeration
    t0 = time.perf_counter()
    ann, entries = L.source_statistics([1] * 8 + [0] * 8, purity=0.99, brightness=0.9, indistinguishability=0.95)
    dt = time.perf_counter() - t0
    assert ann and len(entries) == 26068
    assert sum(p for _, p in entries) == pytest.approx(1.0, abs=1e-12)
    assert dt < 30

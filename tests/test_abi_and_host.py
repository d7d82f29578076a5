"""CPU-side checks: the C-ABI library loads and exports every symbol include/lwb200.h declares, the
host-only Fock indexing is bit-exact with the reference order, and the shard arithmetic is sound."""
import os
import re

import numpy as np
import pytest

import lightworks_b200 as lb
from lightworks_b200 import _lib, sharding
from oracle import oracle as O  # noqa: N812
from tests.golden_util import golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "lwb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lwb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.cdll()
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/lwb200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names
    assert lib.lwb_version() >= 100


def test_compute_calls_fail_loudly_without_gpu():
    try:
        import torch

        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    with pytest.raises(lb.BackendError):
        _lib.Context(0)
    with pytest.raises(lb.BackendError):  # no silent CPU fallback in the product path
        lb.PermanentBackend().probability_amplitude(np.eye(2), [1, 0], [1, 0])
    circ = lb.CompiledCircuitView(np.eye(2))
    with pytest.raises(lb.BackendError):  # nor in the widened rows: sampling and source-statistics combination
        lb.PermanentBackend().sample_outputs(circ, [1, 0], 10, seed=1)
    with pytest.raises(lb.BackendError):
        lb.pdist_calc(circ, {(1, 0): 1.0}, lb.SLOSBackend())
    with pytest.raises(lb.BackendError):  # nor behind the multi-GPU entry points
        lb.Comm.local(2)
    lb.set_n_gpus(2)
    try:
        with pytest.raises(lb.BackendError):
            lb.PermanentBackend().full_probability_distribution(lb.CompiledCircuitView(np.eye(8)), [1] * 7 + [0])
    finally:
        lb.set_n_gpus(1)


def test_keyspace_layout_host():
    """Dense-distribution key space (include/lwb200.h): photon-number blocks side by side, fock_basis order inside."""
    m, kmax = 4, 3
    size = _lib.keyspace_size(m, kmax)
    assert size == sum(lb.fock_count(m, k) for k in range(kmax + 1))
    states = _lib.keyspace_states(m, kmax, np.arange(size))
    expect = [s for k in range(kmax + 1) for s in O.py_fock_basis(m, k)]
    assert states.tolist() == expect
    for i, s in enumerate(expect):
        assert _lib.keyspace_index(s) == i
        assert _lib.keyspace_state(m, kmax, i).tolist() == s
    with pytest.raises(ValueError):
        _lib.keyspace_state(m, kmax, size)


@pytest.mark.parametrize("rec", golden().of_kind("fock"), ids=lambda r: f"m{r['m']}n{r['n']}")
def test_host_rank_unrank_bit_exact(rec):
    ref = golden().arr(rec["basis"])
    m, n = rec["m"], rec["n"]
    assert lb.fock_count(m, n) == len(ref)
    step = max(1, len(ref) // 400)
    for r in list(range(0, len(ref), step)) + [len(ref) - 1]:
        assert np.array_equal(lb.fock_unrank(m, n, r), ref[r])
        assert lb.fock_rank(ref[r]) == r


def test_slos_order_matches_reference_dict_order():
    for rec in golden().of_kind("slos"):
        keys = golden().arr(rec["keys"])
        m, n = keys.shape[1], int(keys[0].sum())
        step = max(1, len(keys) // 300)
        for r in range(0, len(keys), step):
            assert np.array_equal(lb.slos_unrank(m, n, r), keys[r])
            assert lb.slos_rank(keys[r]) == r


def test_rank_formulas_at_scale():
    # C3 (24 modes, 10 photons): rank <-> unrank round trip at random ranks and the extremes
    S = lb.fock_count(24, 10)  # noqa: N806
    assert S == 92561040 and lb.fock_count(40, 20) == 2794563003870330
    rng = np.random.default_rng(0)
    for r in [0, 1, S - 1, *rng.integers(0, S, 200).tolist()]:
        s = lb.fock_unrank(24, 10, int(r))
        assert s.sum() == 10 and lb.fock_rank(s) == r
        t = lb.slos_unrank(24, 10, int(r))
        assert t.sum() == 10 and lb.slos_rank(t) == r
    assert lb.fock_unrank(24, 10, 0).tolist() == [10] + [0] * 23
    assert lb.fock_unrank(24, 10, S - 1).tolist() == [0] * 23 + [10]
    with pytest.raises(ValueError):
        lb.fock_unrank(24, 10, S)
    with pytest.raises(ValueError):
        lb.fock_count(300, 3)
    with pytest.raises(ValueError):
        lb.fock_count(60, 28)  # C(87,28) does not fit 64-bit ranks: refused, not wrapped


def test_shard_ranges_partition_exactly():
    for S in [1, 7, 56, 12376, 92561040]:  # noqa: N806
        for world in [1, 2, 3, 4, 8]:
            rs = [sharding.rank_range(S, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == S
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1


def test_state_mirror_matches_reference_semantics():
    s = lb.State([1, 0, 2])
    assert str(s) == "|1,0,2>" and s.n_photons == 3 and s.n_modes == 3
    assert s == lb.State((1, 0, 2)) and hash(s) == hash(lb.State([1, 0, 2]))
    assert (s + lb.State([0])).s == [1, 0, 2, 0] and s[:2] == lb.State([1, 0]) and s[2] == 2
    c = lb.CompiledCircuitView(np.eye(5), n_modes=3)
    assert c.loss_modes == 2
    assert O.fock_count(5, 2) == lb.fock_count(5, 2)


def test_multiplicity_plan_host():
    """The pre-computed walks of the multiplicity-aware permanent (csrc/k1x.cu), checked without a GPU:
    every partition of n appears once, terms = prod(s_i + 1) / factor, and the binomial coefficients of a walk
    sum to 2^n / factor (each tuple v carries prod C(s_i, v_i) sign vectors of the expanded-row Glynn sum)."""
    from math import prod

    def n_partitions(n):
        p = [1] + [0] * n
        for k in range(1, n + 1):
            for i in range(k, n + 1):
                p[i] += p[i - k]
        return p[n]

    for n in (2, 5, 10, 13):
        plan = _lib.multiplicity_plan(n)
        assert len(plan) == n_partitions(n)
        assert len({tuple(p["s"]) for p in plan}) == len(plan)
        for p in plan:
            s = p["s"]
            assert sum(s) == n and s == sorted(s, reverse=True)
            odd = any(x % 2 for x in s)
            assert p["factor"] == (2 if odd else 1)
            assert p["terms"] == prod(x + 1 for x in s) // p["factor"]
            assert p["abs_coef_sum"] == 2**n / p["factor"]
    with pytest.raises(ValueError):
        _lib.multiplicity_plan(17)


@pytest.mark.parametrize("m,n", [(1, 5), (2, 12), (3, 2), (3, 9), (5, 7), (6, 3), (6, 9), (7, 13), (8, 8), (10, 10),
                                 (12, 6), (16, 8), (20, 6), (30, 5), (40, 4), (64, 3), (128, 2), (24, 8), (16, 12)])
def test_slos_layer_plan_self_check(m, n):
    """lwb_slos_plan_check (host only): every unit of the SLOS layer plan (m modes, n photons) replayed child by child
    with the kernels' own index arithmetic - child rank, the parent of every prefix stream, the parent and the weight
    of every trailing run of the unit kernel and of the record kernel - against exact lexicographic ranks.  The last
    two shapes hold 7.9e6 / 1.7e7 children in 8192-child units with up to 8 trailing photons."""
    info = _lib.slos_plan_check(m, n)
    assert info["children"] == _lib.fock_count(m, n)
    assert info["units"] >= 1 and info["items"] >= 1 and info["items"] <= info["units"]


def test_cumulative_work_is_exact_and_cuts_balance():
    """sharding.cumulative_work against a brute-force sum over fock_basis (terms of the multiplicity-aware walk,
    csrc/k1x.h), and the bisection cuts of work_balanced_ranges."""
    from lightworks_b200 import sharding

    for m, n in [(3, 2), (4, 3), (5, 4), (6, 5), (4, 6), (7, 3)]:
        cum = [0]
        for s in O.py_fock_basis(m, n):
            occ = [v for v in s if v > 0]
            t = int(np.prod([v + 1 for v in occ]))
            if any(v % 2 for v in occ):
                t //= 2
            cum.append(cum[-1] + t)
        for r in range(len(cum)):
            assert sharding.cumulative_work(m, n, r) == cum[r]
    # the walk's own plan (lwb_multiplicity_plan) agrees with the closed form of the total
    m, n = 24, 10
    total = lb.fock_count(m, n)
    assert sharding.cumulative_work(m, n, total) / total == pytest.approx(sharding._average_terms(m, n), rel=1e-12)
    for world in (2, 8, 16):
        ranges = sharding.work_balanced_ranges(m, n, world)
        assert ranges[0][0] == 0 and ranges[-1][1] == total and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        c0 = sharding.STATE_OVERHEAD_TERMS
        w = [sharding.cumulative_work(m, n, b, c0) - sharding.cumulative_work(m, n, a, c0) for a, b in ranges]
        assert max(w) / (sum(w) / world) < 1.0001


def test_native_shard_cuts_equal_the_python_statement():
    """lwb_shard_cuts (csrc/comm.cu, what the communicators use) against lightworks_b200.sharding.work_cuts, the
    readable statement of the same exact-integer arithmetic: identical cuts, including piece fractions."""
    from lightworks_b200 import sharding

    for m, n, world in [(24, 10, 8), (16, 8, 4), (20, 12, 8), (12, 7, 3), (9, 7, 5), (18, 14, 8), (6, 3, 4),
                        (30, 16, 8), (24, 10, 2), (40, 20, 8), (10, 9, 7)]:
        fr = [k / world for k in range(world + 1)]
        assert lb.shard_cuts(m, n, fr) == sharding.work_cuts(m, n, fr), (m, n, world)
    fr = [(r + a) / 8 for r in range(8) for a in (0.0, 0.6, 0.9)] + [1.0]
    assert lb.shard_cuts(24, 10, fr) == sharding.work_cuts(24, 10, fr)
    with pytest.raises(ValueError):
        lb.shard_cuts(200, 10, [0.0, 1.0])


@pytest.mark.reference
def test_reference_source_suite_with_native_statistics():
    """The reference's own tests/emulator/source_test.py with Source._build_statistics answered by
    lwb_source_statistics (lightworks_b200.install(); host only, no GPU involved)."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "oracle.run_reference_tests", "--b200", "emulator/source_test.py"], cwd=root,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    assert " passed" in r.stdout


@pytest.mark.reference
def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's own CPU path, no GPU involved) prints one JSON line with the
    contract keys; under torchrun only rank 0 prints."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-sample", "200"]
    r = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "higher_is_better", "cpu_baseline", "e2e", "config"):
        assert key in line
    assert line["impl"] == "reference" and line["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 0
    r1 = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=600, env={**os.environ, "RANK": "1", "WORLD_SIZE": "2"})
    assert r1.returncode == 0 and r1.stdout.strip() == ""


def test_host_entry_points_reject_bad_arguments():
    """Limits and argument errors of the host-only entry points come back as ValueError with a message, never as a
    wrong answer (include/lwb200.h: LWB_E_ARG / LWB_E_LIMIT)."""
    with pytest.raises(ValueError):
        lb.fock_count(200, 3)  # > 128 modes
    with pytest.raises(ValueError):
        lb.fock_unrank(4, 2, lb.fock_count(4, 2))  # rank outside the space
    with pytest.raises(ValueError):
        _lib.keyspace_size(0, 2)
    with pytest.raises(ValueError):
        _lib.keyspace_states(3, 2, [_lib.keyspace_size(3, 2)])
    with pytest.raises(ValueError):  # 5^12 label classes: beyond the 2^26 enumeration bound, refused up front
        _lib.source_statistics([1] * 12, purity=0.9, brightness=0.9, indistinguishability=0.9)
    with pytest.raises(ValueError):
        _lib.set_option("no_such_option", 1)
    assert "option" in _lib.cdll().lwb_last_error().decode()
    # zero threshold keeps every class; the probabilities still sum to one
    ann, entries = _lib.source_statistics([1, 1, 0], purity=0.8, brightness=0.7, indistinguishability=0.6,
                                          probability_threshold=0.0)
    assert ann and sum(p for _, p in entries) == pytest.approx(1.0, abs=1e-13)
    assert len(entries) > len(_lib.source_statistics([1, 1, 0], purity=0.8, brightness=0.7, indistinguishability=0.6,
                                                     probability_threshold=5e-2)[1])

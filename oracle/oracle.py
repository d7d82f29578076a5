"""
oracle.oracle -- Python face of the CPU oracle.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by lightworks_b200.

Two layers:
  * `py_*` functions: literal pure-Python restatements of the reference functions
    (small cases only), each citing the reference file:line it follows
    (paths relative to /root/reference).
  * ctypes bindings to liblw_oracle.so (oracle/lw_oracle.c), the same algorithms in C
    for sizes Python cannot reach.  `build()` compiles it with gcc.
Both are pinned against the reference's literals and against the imported reference
(tests/test_oracle_golden.py, tests/golden/).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from math import factorial, prod

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liblw_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "lw_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p, f64p, vp = C.POINTER(C.c_uint8), C.POINTER(C.c_double), C.c_void_p
        L.lwo_fock_count.restype = C.c_uint64
        L.lwo_fock_count.argtypes = [C.c_int, C.c_int]
        L.lwo_fock_basis.restype = C.c_uint64
        L.lwo_fock_basis.argtypes = [C.c_int, C.c_int, vp, C.c_uint64]
        L.lwo_perm.restype = None
        L.lwo_perm.argtypes = [vp, C.c_int, vp]
        L.lwo_perm_hp.restype = None
        L.lwo_perm_hp.argtypes = [vp, C.c_int, vp, C.c_int]
        L.lwo_perm_batch.restype = None
        L.lwo_perm_batch.argtypes = [vp, C.c_int, C.c_int, vp]
        L.lwo_perm_batch_mt.restype = None
        L.lwo_perm_batch_mt.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int]
        L.lwo_probability_amplitude.restype = C.c_int
        L.lwo_probability_amplitude.argtypes = [vp, C.c_int, vp, vp, vp]
        L.lwo_amplitude_matrix.restype = C.c_int
        L.lwo_amplitude_matrix.argtypes = [vp, C.c_int, vp, C.c_int, vp, C.c_int, vp]
        L.lwo_basis_probabilities.restype = C.c_int
        L.lwo_basis_probabilities.argtypes = [vp, C.c_int, vp, C.c_uint64, C.c_uint64, vp]
        L.lwo_basis_probabilities_mt.restype = C.c_int
        L.lwo_basis_probabilities_mt.argtypes = [vp, C.c_int, vp, C.c_uint64, C.c_uint64, vp, C.c_int]
        L.lwo_pdist_permanent.restype = C.c_int64
        L.lwo_pdist_permanent.argtypes = [vp, C.c_int, C.c_int, vp, C.c_double, vp, vp, C.c_int64]
        L.lwo_slos_calculate.restype = C.c_int64
        L.lwo_slos_calculate.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int64]
        L.lwo_pdist_slos.restype = C.c_int64
        L.lwo_pdist_slos.argtypes = [vp, C.c_int, C.c_int, vp, C.c_double, vp, vp, C.c_int64]
        _lib = L
    return _lib


def _c128(a):
    return np.ascontiguousarray(a, dtype=np.complex128)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# --------------------------------------------------------------------------- #
# Input generator: lightworks/sdk/utils/random.py:45-67 (random_unitary)       #
# pinned by tests/sdk/util_test.py:53-69 (seed 111)                            #
# --------------------------------------------------------------------------- #
def random_unitary(n: int, seed: int | None = None) -> np.ndarray:
    from scipy.stats import unitary_group

    return unitary_group.rvs(n, random_state=seed)


# --------------------------------------------------------------------------- #
# Pure-Python literal restatements (small sizes)                               #
# --------------------------------------------------------------------------- #
def py_fock_basis(n_modes: int, n_photons: int) -> list[list[int]]:
    """lightworks/emulator/utils/state.py:22-35"""

    def _sums(length, total):
        if length == 1:
            yield [total]
        else:
            for value in range(total + 1):
                for perm_ in _sums(length - 1, total - value):
                    yield [*perm_, value]

    return list(_sums(n_modes, n_photons))


def py_partition(unitary, in_state, out_state):
    """lightworks/emulator/backends/permanent.py:166-178"""
    n_modes = len(in_state)
    x = [i for i in range(n_modes) for _ in range(out_state[i])]
    y = [i for i in range(n_modes) for _ in range(in_state[i])]
    return np.asarray(unitary)[np.ix_(x, y)]


def py_perm(A) -> complex:  # noqa: N803
    """thewalrus==0.20.0 perm / perm_bbfg, restated (see lw_oracle.c)."""
    A = np.asarray(A, dtype=np.complex128)  # noqa: N806
    n = A.shape[0]
    if n == 0:
        return 1.0 + 0j
    if n == 1:
        return complex(A[0, 0])
    if n == 2:
        return complex(A[0, 0] * A[1, 1] + A[0, 1] * A[1, 0])
    if n == 3:
        return complex(
            A[0, 2] * A[1, 1] * A[2, 0]
            + A[0, 1] * A[1, 2] * A[2, 0]
            + A[0, 2] * A[1, 0] * A[2, 1]
            + A[0, 0] * A[1, 2] * A[2, 1]
            + A[0, 1] * A[1, 0] * A[2, 2]
            + A[0, 0] * A[1, 1] * A[2, 2]
        )
    row_comb = A.sum(axis=0)
    total = 0j
    old_gray = 0
    sign = 1
    num_loops = 2 ** (n - 1)
    for bin_index in range(1, num_loops + 1):
        total += sign * np.prod(row_comb)
        new_gray = bin_index ^ (bin_index // 2)
        k = (old_gray ^ new_gray).bit_length() - 1
        direction = 2 * ((old_gray > new_gray) - (old_gray < new_gray))
        row_comb = row_comb + A[k] * direction
        old_gray = new_gray
        sign = -sign
    return complex(total / num_loops)


def py_perm_naive(A) -> complex:  # noqa: N803
    """Definition of the permanent (sum over permutations); independent cross-check."""
    from itertools import permutations

    A = np.asarray(A, dtype=np.complex128)  # noqa: N806
    n = A.shape[0]
    return complex(sum(prod(A[i, s[i]] for i in range(n)) for s in permutations(range(n))))


def py_probability_amplitude(unitary, input_state, output_state) -> complex:
    """lightworks/emulator/backends/permanent.py:53-83"""
    factor_m = prod([factorial(i) for i in input_state])
    factor_n = prod([factorial(i) for i in output_state])
    return py_perm(py_partition(unitary, input_state, output_state)) / np.sqrt(factor_m * factor_n)


def py_pdist_permanent(U_full, n_modes, input_state, threshold=1e-9):  # noqa: N803
    """permanent.py:116-163; returns dict[tuple, float] in insertion order."""
    input_state = list(input_state)
    if sum(input_state) == 0:
        return {tuple([0] * n_modes): 1.0}
    loss_modes = U_full.shape[0] - n_modes
    pdist: dict[tuple, float] = {}
    input_state = input_state + [0] * loss_modes
    for ostate in py_fock_basis(len(input_state), sum(input_state)):
        if sum(ostate[:n_modes]) == 0:
            continue
        p = abs(py_probability_amplitude(U_full, input_state, ostate)) ** 2
        if p > threshold:
            key = tuple(ostate[:n_modes])
            if key in pdist:
                pdist[key] += p
            else:
                pdist[key] = p
    total = sum(pdist.values())
    if total < 1 and loss_modes > 0:
        pdist[tuple([0] * n_modes)] = 1 - total
    return pdist


def py_slos_calculate(unitary, input_state) -> dict[tuple, complex]:
    """lightworks/emulator/backends/slos.py:82-138 (calculate, a_i_dagger, add_dicts)"""
    input_state = list(input_state)
    p = [m for m, n in enumerate(input_state) for _ in range(n)]
    n_modes = unitary.shape[0]
    m = 1 / np.sqrt(int(np.prod([factorial(i) for i in input_state])))
    input_ = {tuple(n_modes * [0]): m}
    for i in p:
        output: dict[tuple, complex] = {}
        for j in range(n_modes):
            mult = unitary[j, i]
            for key, value in input_.items():
                k = list(key)
                k[j] += 1
                v = k[j] ** 0.5 * value * mult
                tk = tuple(k)
                if tk in output:
                    output[tk] += v
                else:
                    output[tk] = v
        input_ = output
    return input_


def py_pdist_slos(U_full, n_modes, input_state, threshold=1e-9):  # noqa: N803
    """slos.py:43-80"""
    input_state = list(input_state)
    if sum(input_state) == 0:
        return {tuple([0] * n_modes): 1.0}
    loss_modes = U_full.shape[0] - n_modes
    input_state = input_state + [0] * loss_modes
    full = py_slos_calculate(U_full, input_state)
    pdist: dict[tuple, float] = {}
    for s, p in full.items():
        if abs(p) ** 2 > threshold:
            key = tuple(s[:n_modes])
            if key in pdist:
                pdist[key] += abs(p) ** 2
            else:
                pdist[key] = abs(p) ** 2
    return pdist


# --------------------------------------------------------------------------- #
# C-backed versions                                                            #
# --------------------------------------------------------------------------- #
def fock_count(m: int, n: int) -> int:
    return int(lib().lwo_fock_count(m, n))


def fock_basis(m: int, n: int, limit: int | None = None) -> np.ndarray:
    S = fock_count(m, n)  # noqa: N806
    cap = S if limit is None else min(S, limit)
    out = np.zeros((cap, m), dtype=np.uint8)
    lib().lwo_fock_basis(m, n, _p(out), cap)
    return out


def perm(A) -> complex:  # noqa: N803
    A = _c128(A)  # noqa: N806
    out = np.zeros(2)
    lib().lwo_perm(_p(A), A.shape[0], _p(out))
    return complex(out[0], out[1])


def perm_hp(A, threads: int | None = None) -> complex:  # noqa: N803
    """Extended-precision ADJUDICATOR (lw_oracle.c lwo_perm_hp): the same Glynn sum in long double with compensated
    accumulation.  Not a restatement of the reference: the yardstick that tells which side of a mismatch at n >= 20
    carries the rounding error (the reference's incremental double-precision loop does)."""
    A = _c128(A)  # noqa: N806
    out = np.zeros(2)
    lib().lwo_perm_hp(_p(A), A.shape[0], _p(out), threads or (os.cpu_count() or 1))
    return complex(out[0], out[1])


def perm_batch(A, threads: int = 1) -> np.ndarray:  # noqa: N803
    A = _c128(A)  # noqa: N806
    B, n = A.shape[0], A.shape[1]  # noqa: N806
    out = np.zeros(B, dtype=np.complex128)
    if threads > 1:
        lib().lwo_perm_batch_mt(_p(A), B, n, _p(out), threads)
    else:
        lib().lwo_perm_batch(_p(A), B, n, _p(out))
    return out


def probability_amplitude(U, in_state, out_state) -> complex:  # noqa: N803
    U = _c128(U)  # noqa: N806
    i, o = _u8(in_state), _u8(out_state)
    out = np.zeros(2)
    if lib().lwo_probability_amplitude(_p(U), U.shape[0], _p(i), _p(o), _p(out)):
        raise ValueError("photon number mismatch")
    return complex(out[0], out[1])


def amplitude_matrix(U, ins, outs) -> np.ndarray:  # noqa: N803
    U = _c128(U)  # noqa: N806
    ins, outs = _u8(ins), _u8(outs)
    amps = np.zeros((ins.shape[0], outs.shape[0]), dtype=np.complex128)
    if lib().lwo_amplitude_matrix(_p(U), U.shape[0], _p(ins), ins.shape[0], _p(outs), outs.shape[0], _p(amps)):
        raise ValueError("photon number mismatch")
    return amps


def basis_probabilities(U, in_state, r0=0, cnt=None, threads: int = 1) -> np.ndarray:  # noqa: N803
    U = _c128(U)  # noqa: N806
    i = _u8(in_state)
    m = U.shape[0]
    S = fock_count(m, int(i.sum()))  # noqa: N806
    cnt = S - r0 if cnt is None else cnt
    probs = np.zeros(cnt)
    if threads > 1:
        rc = lib().lwo_basis_probabilities_mt(_p(U), m, _p(i), r0, cnt, _p(probs), threads)
    else:
        rc = lib().lwo_basis_probabilities(_p(U), m, _p(i), r0, cnt, _p(probs))
    if rc:
        raise ValueError("bad rank range")
    return probs


def _pdist(fn, U_full, n_modes, in_state, threshold):  # noqa: N803
    U_full = _c128(U_full)  # noqa: N806
    i = _u8(in_state)
    m_tot = U_full.shape[0]
    cap = fock_count(m_tot, int(i.sum())) + 2
    keys = np.zeros((cap, n_modes), dtype=np.uint8)
    vals = np.zeros(cap)
    n = fn(_p(U_full), m_tot, n_modes, _p(i), threshold, _p(keys), _p(vals), cap)
    if n < 0:
        raise RuntimeError("oracle pdist failed")
    return keys[:n].copy(), vals[:n].copy()


def pdist_permanent(U_full, n_modes, in_state, threshold=1e-9):  # noqa: N803
    """(keys uint8[K,n_modes], vals float64[K]) in dict insertion order; permanent.py:116-163"""
    return _pdist(lib().lwo_pdist_permanent, U_full, n_modes, in_state, threshold)


def pdist_slos(U_full, n_modes, in_state, threshold=1e-9):  # noqa: N803
    """(keys, vals) in dict insertion order; slos.py:43-80"""
    return _pdist(lib().lwo_pdist_slos, U_full, n_modes, in_state, threshold)


def slos_calculate(U, in_state):  # noqa: N803
    """(keys uint8[S,m], amps complex128[S]) in dict insertion order; slos.py:82-105"""
    U = _c128(U)  # noqa: N806
    i = _u8(in_state)
    m = U.shape[0]
    S = fock_count(m, int(i.sum()))  # noqa: N806
    keys = np.zeros((S, m), dtype=np.uint8)
    amps = np.zeros(S, dtype=np.complex128)
    n = lib().lwo_slos_calculate(_p(U), m, _p(i), _p(keys), _p(amps), S)
    if n < 0:
        raise RuntimeError("oracle slos failed")
    return keys[:n], amps[:n]


# ---- sampling ---------------------------------------------------------
def sample_indices(probs, uniform, threshold: float = -1.0) -> np.ndarray:
    """Positions drawn by `uniform` from the distribution `probs`: restates
    SamplerRunner._gen_samples_from_dist (simulation/sampler.py:329-349),
        probs /= sum(probs); rng.choice(range(len(probs)), p=probs, size=N),
    through numpy's published algorithm for Generator.choice with p (numpy/random/_generator.pyx):
        cdf = p.cumsum(); cdf /= cdf[-1]; idx = cdf.searchsorted(rng.random(N), side="right").
    Entries with p <= threshold count as zero: the reference's distribution dict does not hold them
    (permanent.py:151, slos.py:74) and a zero-width interval is never drawn.  Pinned against numpy itself
    (tests/test_oracle_golden.py::test_sample_indices_is_numpy_choice)."""
    p = np.asarray(probs, dtype=np.float64).copy()
    p[~(p > threshold)] = 0.0
    cdf = p.cumsum()
    cdf /= cdf[-1]
    return cdf.searchsorted(np.asarray(uniform, dtype=np.float64), side="right").astype(np.uint64)


def sample_states(dist: dict, n_samples: int, seed=None) -> dict:
    """_gen_samples_from_dist(dist, n_samples, seed, normalize=True) verbatim in behaviour (sampler.py:329-349):
    {state: count} for a dict distribution, drawn with numpy's Generator exactly as the reference does."""
    from collections import Counter

    mapping = dict(enumerate(dist))
    probs = np.array(list(dist.values()), dtype=float)
    probs /= sum(probs)
    rng = np.random.default_rng(seed)
    samples = rng.choice(range(len(mapping)), p=probs, size=n_samples)
    return {mapping[n]: c for n, c in Counter(samples).items()}


# ---- source-statistics combination -------------------------------------
def py_pdist_calc(circuit, inputs: dict, probability_func) -> dict:
    """pdist_calc (simulation/probability_distribution.py:25-73) for State inputs, with tuples for states.
    `circuit` needs n_modes and loss_modes; probability_func(circuit, state_tuple) -> dict[tuple, float]."""
    pdist: dict = {}
    for istate, prob in inputs.items():                      # :58
        sub_dist = probability_func(circuit, istate)         # :60
        if not pdist:                                        # :61-65
            pdist = dict(sub_dist) if prob == 1 else {s: p * prob for s, p in sub_dist.items()}
        else:                                                # :66-71
            for s, p in sub_dist.items():
                if s in pdist:
                    pdist[s] += p * prob
                else:
                    pdist[s] = p * prob
    total_prob = sum(pdist.values())                         # :73
    if total_prob < 1 and circuit.loss_modes > 0:            # :74-75
        pdist[tuple([0] * circuit.n_modes)] = 1 - total_prob
    return pdist


def py_annotated_pdist_calc(circuit, inputs: dict, probability_func) -> dict:
    """annotated_state_pdist_calc (simulation/probability_distribution.py:78-164).  Annotated states are tuples
    of per-mode label tuples; State.merge (sdk/state/state.py) is the mode-wise sum of occupations."""
    n_modes = circuit.n_modes
    unique_inputs: list = []
    input_combinations: dict = {}
    for state in inputs:                                     # :105-121
        all_labels = [lab for mode in state for lab in mode]
        if all_labels:
            results = {lab: [0] * n_modes for lab in all_labels}
            for i, mode in enumerate(state):
                for m in mode:
                    results[m][i] += 1
            states = [tuple(s) for s in results.values()]
        else:
            states = [tuple([0] * n_modes)]
        for s in states:
            if s not in unique_inputs:
                unique_inputs.append(s)
        input_combinations[state] = states
    unique_results = {s[:n_modes]: list(probability_func(circuit, s).items()) for s in unique_inputs}  # :124-133
    stats_dict: dict = {}
    for istate, combination in input_combinations.items():   # :137
        pdist: dict = {}
        for d in combination:                                # :140
            if not pdist:
                pdist = dict(unique_results[d])
            else:
                new_pdist: dict = {}
                for output, p1 in pdist.items():             # :147-156
                    for result, p2 in unique_results[d]:
                        new_state = tuple(a + b for a, b in zip(output, result))
                        if new_state not in new_pdist:
                            new_pdist[new_state] = p1 * p2
                        else:
                            new_pdist[new_state] += p1 * p2
                pdist = new_pdist
        ip = inputs[istate]                                  # :159-164
        for ostate, op in pdist.items():
            if ostate not in stats_dict:
                stats_dict[ostate] = ip * op
            else:
                stats_dict[ostate] += ip * op
    return stats_dict


def pdist_func(backend: str = "permanent", threshold: float = 1e-9):
    """probability_func for the two restatements above, from the oracle's own full distributions."""
    fn = pdist_permanent if backend == "permanent" else pdist_slos

    def f(circuit, state):
        keys, vals = fn(circuit.U_full, circuit.n_modes, list(state)[: circuit.n_modes], threshold)
        return {tuple(int(v) for v in k): float(p) for k, p in zip(keys, vals)}

    return f


# ---- imperfect-source input statistics (second half) ---------------------
def py_build_statistics(state, purity=1.0, brightness=1.0, indistinguishability=1.0, probability_threshold=1e-9):
    """Source._build_statistics (emulator/components/source.py:134-161) restated with tuples: returns
    (annotated, dict) in the reference's dict order.  States are occupation tuples (basic method, :163-224) or
    tuples of per-mode label tuples (full method, :226-391); thresholding and renormalisation as :394-406."""
    state = tuple(int(v) for v in state)
    n_modes = len(state)
    if purity == 1 and indistinguishability == 1:                         # :155-157
        stats: list = []
        for mode, count in enumerate(state):                              # :177-209
            for _ in range(count):
                sub_s = [(brightness, tuple(1 if mode == j else 0 for j in range(n_modes)))]
                if brightness < 1:
                    sub_s.append((1 - brightness, tuple([0] * n_modes)))
                if not stats:
                    stats = sub_s
                else:
                    stats = [(p1 * p2, tuple(x + y for x, y in zip(s1, s2))) for p1, s1 in stats for p2, s2 in sub_s]
        dist: dict = {}
        for p, s in stats:                                                # :211-218
            dist[s] = dist.get(s, 0) + p if s in dist else p
        if not dist:                                                      # :221-223
            dist[state] = 1.0
        annotated = False
    else:
        if purity <= 0.5:                                                 # :442-443
            raise ValueError("Purity values below 0.5 not supported.")
        if purity < 1:                                                    # :444-447
            g2 = 1 - purity
            b = 2 * (1 - (1 / g2))
            p1 = 1 - (-b - (b**2 - 4) ** 0.5) / 2
        else:
            p1 = 1
        counter = [1]

        def single_photon():                                              # :353-391
            nu, p_i = brightness, indistinguishability**0.5
            p_d, p2 = 1 - p_i, 1 - p1
            c0 = 1 - nu * (p1 + p2 * nu + 2 * (1 - nu) * p2)
            c1 = p_i * nu * (p1 + (1 - nu) * p2)
            c1d = p_d * nu * (p1 + (1 - nu) * p2)
            c1dp = nu * (1 - nu) * p2
            c12d = nu**2 * p_i * p2
            c1d2d = nu**2 * p_d * p2
            dpc, mpc = counter[0], counter[0] + 1
            counter[0] += 2
            to_add = [([], c0), ([0], c1), ([dpc], c1d), ([mpc], c1dp), ([0, mpc], c12d), ([dpc, mpc], c1d2d)]
            return [(s, p) for s, p in to_add if p > 0]

        def single_mode(n):                                               # :318-351
            if n == 0:
                return {((),): 1.0}
            mode_dist: list = []
            for _ in range(n):
                calc = single_photon()
                mode_dist = calc if not mode_dist else [(d1[0] + d2[0], d1[1] * d2[1]) for d1 in mode_dist for d2 in calc]
            out: dict = {}
            for s, p in mode_dist:
                key = (tuple(sorted(s)),)
                out[key] = out[key] + p if key in out else p
            return out

        full: dict = {}
        for n in state:                                                   # :292-316 (empty-mode grouping is an optimisation)
            calc = single_mode(n)
            full = calc if not full else {s1 + s2: p1_ * p2_ for s1, p1_ in full.items() for s2, p2_ in calc.items()}
        dist = {}
        for s, p in full.items():                                         # :252-283
            labels = list(dict.fromkeys(lab for mode in s for lab in mode))
            mapping = {lab: i for i, lab in enumerate(labels)}
            key = tuple(tuple(sorted(mapping[lab] for lab in mode)) for mode in s)
            dist[key] = dist[key] + p if key in dist else p
        annotated = True
    kept = {s: p for s, p in dist.items() if p >= probability_threshold}   # :401-406
    total = sum(kept.values())
    return annotated, {s: p / total for s, p in kept.items()}

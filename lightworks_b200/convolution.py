"""
pdist_calc / annotated_state_pdist_calc on the GPU.

Mirror of lightworks/emulator/simulation/probability_distribution.py:25-164 — same name, arguments and meaning —
with the per-input distributions kept on the device as dense vectors over the key space K(n_modes, n_photons)
(include/lwb200.h) and the nested `State.merge` dict loops (:150-158) replaced by `lwb_dist_convolve`.

    pdist_calc(circuit, inputs, backend) -> LazyDistribution (a Mapping[State, float])

`inputs` is what `Source._build_statistics` returns: dict[State, p] (perfect source / brightness only) or
dict[AnnotatedState, p] (labels = mutually distinguishable photon groups).  `backend` is a lightworks_b200 backend
object (its `full_probability_distribution` is what the reference would call back).

Differences from the reference, both outside its numerical contract:
  * entries come back in key-space order (photon number, then fock_basis order), not in first-merged order, so a
    Sampler seeded identically draws different (equally distributed) samples;
  * the products of a chain are formed in a canonical order so partial products are shared between annotated inputs
    (floating-point sums differ at the 1e-16 level).
There is no CPU path: every distribution is computed and combined by liblwb200.so.
"""
from __future__ import annotations

import numpy as np

from . import _lib


def _labels_to_states(annotated, n_modes: int) -> list[tuple[int, ...]]:
    """probability_distribution.py:106-121: one occupation vector per label, in first-seen label order."""
    results: dict = {}
    for i, mode in enumerate(annotated):
        for lab in mode:
            results.setdefault(lab, [0] * n_modes)[i] += 1
    if not results:
        return [tuple([0] * n_modes)]
    return [tuple(v) for v in results.values()]


def _is_annotated(state) -> bool:
    s = state.s if hasattr(state, "s") else state
    return len(s) > 0 and isinstance(s[0], (list, tuple))


class SourceClasses:
    """The inputs of an imperfect source in array form (lwb_source_statistics): per input the occupation vector of the
    mutually indistinguishable photons and of the photons that carry a label of their own.  Accepted by pdist_calc in
    place of the reference's dict[AnnotatedState, p]; skips building and re-parsing label lists."""

    def __init__(self, annotated: bool, group_occ: np.ndarray, single_occ: np.ndarray, probs: np.ndarray):
        self.annotated, self.group_occ, self.single_occ, self.probs = annotated, group_occ, single_occ, probs

    @classmethod
    def from_source(cls, state, purity=1.0, brightness=1.0, indistinguishability=1.0, probability_threshold=1e-9):
        return cls(*_lib.source_classes(state, purity, brightness, indistinguishability, probability_threshold))

    def __len__(self):
        return len(self.probs)

    def combos(self, n_modes: int) -> list:
        """[(list of per-label occupation tuples, p)]: the group (if any photons) and one unit vector per single."""
        if self.group_occ.shape[1] != n_modes:
            raise ValueError("source statistics and circuit have different mode counts")
        eye = [tuple(1 if i == j else 0 for i in range(n_modes)) for j in range(n_modes)]
        out = []
        for g, b, p in zip(self.group_occ.tolist(), self.single_occ.tolist(), self.probs.tolist()):
            states = [tuple(g)] if any(g) or not any(b) else []
            for j, c in enumerate(b):
                states += [eye[j]] * c
            out.append((states, p))
        return out


class DistributionAlgebra:
    """Device-resident per-input distributions of one circuit and cached partial products."""

    def __init__(self, ctx, circuit, backend_id, threshold: float, dtype=_lib.LWB_C128, max_cached: int = 4096,
                 max_cached_bytes: int = 8 << 30):
        self.ctx, self.circuit = ctx, circuit
        self.backend_id, self.threshold, self.dtype = backend_id, threshold, dtype
        self._unique: dict[tuple, _lib.DeviceDistribution] = {}
        self._products: dict[tuple, _lib.DeviceDistribution] = {}
        self._max_cached, self._max_bytes, self._bytes = max_cached, max_cached_bytes, 0
        self.n_convolutions = 0
        self._vacuum = None

    def prefetch(self, states) -> None:
        """Every unique input of the source in ONE call (lwb_dists_from_circuit): no per-input synchronisation, and
        with install(n_gpus=N) / set_n_gpus(N) the inputs are computed round-robin on the N GPUs."""
        from .backends import local_comm

        todo = [s for s in dict.fromkeys(states) if s not in self._unique]
        if not todo:
            return
        comm = local_comm()
        if comm is not None and comm.context(0) is not self.ctx:
            comm = None  # the slots must live where the convolutions run
        dists = self.ctx.dists_from_circuit(self.circuit.U_full, self.circuit.n_modes, [list(s) for s in todo],
                                            self.threshold, self.backend_id, self.dtype, comm=comm)
        self._unique.update(zip(todo, dists))

    def unique(self, state: tuple) -> _lib.DeviceDistribution:
        """probability_func(circuit, in_state) for one unique input (probability_distribution.py:124-129)."""
        d = self._unique.get(state)
        if d is None:
            d = self.ctx.dist_from_circuit(self.circuit.U_full, self.circuit.n_modes, list(state), self.threshold,
                                           self.backend_id, self.dtype)
            self._unique[state] = d
        return d

    def vacuum(self) -> _lib.DeviceDistribution:
        """The distribution of "no photons": all weight on the vacuum key."""
        if self._vacuum is None:
            self._vacuum = self.ctx.dist_from_dense(self.circuit.n_modes, 0, np.ones(1))
        return self._vacuum

    def product(self, states: list[tuple]) -> _lib.DeviceDistribution:
        """Convolution of the distributions of `states` (probability_distribution.py:139-158); commutative, so the
        chain is formed in sorted order and every prefix is cached for the other annotated inputs."""
        # single (fully distinguishable) photons first: their classical products are shared by many annotated
        # inputs and stay small; the big indistinguishable group is convolved in last
        chain = tuple(sorted(states, key=lambda s: (sum(s), s)))
        d = self._products.get(chain)
        if d is not None:
            return d
        if len(chain) == 1:
            return self.unique(chain[0])
        d = self.product(list(chain[:-1])) @ self.unique(chain[-1])
        self.n_convolutions += 1
        self._bytes += 8 * d.size
        while self._products and (len(self._products) >= self._max_cached or self._bytes > self._max_bytes):
            old = self._products.pop(next(iter(self._products)))  # oldest first; recomputed if needed again
            self._bytes -= 8 * old.size
        self._products[chain] = d
        return d


def pdist_calc(circuit, inputs, backend, keep_on_device: bool = False):
    """probability_distribution.py:25-73 (State inputs) and :78-164 (AnnotatedState inputs) on the device.
    keep_on_device: return the accumulated DeviceDistribution (dense over K(n_modes, max photons)) instead of
    downloading it; not available where the reference patches the vacuum entry afterwards (State inputs on a
    lossy circuit, :70-73)."""
    from .backends import LazyDistribution, _threshold

    ctx = backend.ctx
    n_modes = circuit.n_modes
    alg = DistributionAlgebra(ctx, circuit, backend._backend_id, _threshold(), backend._dtype)
    if not isinstance(inputs, SourceClasses):
        inputs = dict(inputs)
    if not len(inputs):
        return LazyDistribution(np.zeros((0, n_modes), dtype=np.uint8), np.zeros(0))
    if isinstance(inputs, SourceClasses):
        annotated, combos = inputs.annotated, inputs.combos(n_modes)
    else:
        annotated = _is_annotated(next(iter(inputs)))
        combos = []
        for istate, prob in inputs.items():
            if annotated:
                combos.append((_labels_to_states(istate, n_modes), float(prob)))
            else:
                s = tuple(istate.s if hasattr(istate, "s") else istate)
                combos.append(([s[:n_modes] if len(s) > n_modes else s], float(prob)))
    kmax = max(sum(sum(s) for s in states) for states, _ in combos)
    alg.prefetch(st for states, _ in combos for st in states)
    acc = ctx.dist_zeros(n_modes, kmax)
    # sum_inputs p * D_G (*) prod(rest) = sum_G D_G (*) ( sum_rest p * prod(rest) ) by linearity: the inputs are grouped
    # by their largest photon group G (for a Source: the mutually indistinguishable photons), the cheap products of
    # the remaining groups (single distinguishable photons) are accumulated first, and only one large convolution
    # per distinct G is left (:137-164 run once per annotated input in the reference)
    by_group: dict[tuple, list] = {}
    for states, prob in combos:
        big = max(states, key=lambda st: (sum(st), st))
        rest = list(states)
        rest.remove(big)
        by_group.setdefault(big, []).append((rest, prob))
    for big, items in by_group.items():
        if all(not rest for rest, _ in items):
            acc.add_scaled(sum(p for _, p in items), alg.unique(big))  # :160-164 / :62-69
            continue
        k_rest = max(sum(sum(st) for st in rest) for rest, _ in items)
        others = ctx.dist_zeros(n_modes, k_rest)
        # one launch adds every annotated input of the group, in input order (the same operations as one axpy each)
        others.add_scaled_many([prob for _, prob in items], [alg.product(rest) if rest else alg.vacuum() for rest, _ in items])
        acc.add_scaled(1.0, alg.unique(big) @ others)
        alg.n_convolutions += 1
    if keep_on_device:
        if not annotated and getattr(circuit, "loss_modes", 0) > 0:
            raise ValueError("keep_on_device: the vacuum rule of probability_distribution.py:70-73 needs the host copy")
        acc.n_convolutions = alg.n_convolutions
        return acc
    dense = acc.to_numpy()
    if not annotated and getattr(circuit, "loss_modes", 0) > 0:
        # :70-73: `if total_prob < 1: pdist[vacuum] = 1 - total_prob` OVERWRITES an existing vacuum entry.  When the
        # sub-distributions already carry their vacuum terms the total is 1 up to rounding, and whether the reference
        # keeps the true value or replaces it by ~1e-16 depends on its summation order; a deficit
        # at rounding level is therefore taken as "total == 1" here.
        total = float(dense.sum())
        if 1 - total > 16 * np.finfo(float).eps:
            dense[0] = 1 - total
    idx = np.nonzero(dense)[0]
    keys = _lib.keyspace_states(n_modes, kmax, idx)
    out = LazyDistribution(keys, dense[idx])
    out.n_convolutions = alg.n_convolutions
    return out



def detect(rows: np.ndarray, rng, efficiency: float = 1.0, p_dark: float = 0.0, photon_counting: bool = True) -> np.ndarray:
    """Detector._get_output (components/detector.py:101-135) for every row of `rows` (uint8 [samples, modes]) at
    once: each photon survives with probability `efficiency`, each mode gains one dark count with probability
    p_dark, threshold detectors report min(count, 1).  Only occupied entries draw a variate (most of a Fock state is
    empty), and a single photon is a Bernoulli draw instead of a binomial one."""
    rows = np.ascontiguousarray(rows, dtype=np.uint8)
    if efficiency < 1:
        flat = rows.reshape(-1)
        nz = np.flatnonzero(flat)
        vals = flat[nz]
        kept = (rng.random(len(nz)) < efficiency).astype(np.uint8)
        many = np.flatnonzero(vals > 1)
        if len(many):
            kept[many] = rng.binomial(vals[many], efficiency)
        flat[nz] = kept
    if p_dark > 0:
        rows = rows + (rng.random(rows.shape) < p_dark).astype(np.uint8)
    if not photon_counting:
        rows = np.minimum(rows, 1)
    return rows


def count_rows_first_seen(rows: np.ndarray):
    """Counter(rows) as arrays: (index of the first occurrence of every distinct row, its count), in first-seen order
    like the dict a SamplingResult is built from.  Rows of <= 16 small counts are packed into one 64-bit key (a plain
    integer sort); anything else is compared as byte strings."""
    rows = np.ascontiguousarray(rows, dtype=np.uint8)
    n, m = rows.shape
    if n == 0:
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    bits = max(int(rows.max()).bit_length(), 1)
    if m * bits <= 64:
        key = np.zeros(n, dtype=np.uint64)
        for j in range(m):
            key = (key << np.uint64(bits)) | rows[:, j].astype(np.uint64)
    else:
        key = rows.view(np.dtype((np.void, m))).reshape(-1)
    _, first, counts = np.unique(key, return_index=True, return_counts=True)
    order = np.argsort(first, kind="stable")
    return first[order], counts[order]

def sample_source_outputs(circuit, input_state, n_samples: int, backend, purity: float = 1.0, brightness: float = 1.0,
                          indistinguishability: float = 1.0, probability_threshold: float = 1e-9,
                          detector_efficiency: float = 1.0, p_dark: float = 0.0, photon_counting: bool = True, seed=None,
                          as_arrays: bool = False):
    """`Sampler(circuit, input_state, n_samples, source=Source(...), detector=Detector(...), sampling_mode="input")`
    (simulation/sampler.py:144-237) for circuits without heralds or post-selection, with the output distribution of the
    imperfect source built, kept and sampled on the device:

      Source._build_statistics            -> lwb_source_statistics (host, components/source.py:134-406)
      pdist_calc                          -> dense distributions + merge-convolution (probability_distribution.py:25-164)
      _gen_samples_from_dist              -> lwb_dist_sample (sampler.py:329-349; uniforms from numpy's default_rng(seed))
      Detector._get_output per sample     -> vectorised: binomial thinning with `efficiency`, one dark count per mode
                                             with probability p_dark, threshold detection (components/detector.py:101-135)

    Returns {State: count} like SamplingResult (as_arrays=True: (states uint8[K, m], counts[K]) without building State
    objects).  The drawn output states follow the reference's distribution; the
    detector draws use numpy's generator instead of Python's `random`, so counts are statistically, not bitwise, equal to
    the reference's for a seed."""
    from .backends import _as_list, _state_cls

    ins = _as_list(input_state)
    n_modes = circuit.n_modes
    if len(ins) != n_modes:
        raise ValueError("input state length does not match the circuit")
    classes = SourceClasses.from_source(ins, purity, brightness, indistinguishability, probability_threshold)
    acc = pdist_calc(circuit, classes, backend, keep_on_device=True)
    rng = np.random.default_rng(seed)
    idx, _total = acc.sample(rng.random(int(n_samples)))
    uniq, inverse = np.unique(idx, return_inverse=True)
    out = _lib.keyspace_states(n_modes, acc.kmax, uniq).astype(np.uint8)[inverse.reshape(-1)]  # one row per sample
    rows = detect(out, rng, detector_efficiency, p_dark, photon_counting)
    first, counts = count_rows_first_seen(rows)
    if as_arrays:
        return rows[first], counts
    return {_state_cls(k): int(c) for k, c in zip(rows[first].tolist(), counts.tolist())}

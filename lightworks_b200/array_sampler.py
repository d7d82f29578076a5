"""
Array-native Sampler path behind the reference's operator API.

`FockBackend.run_sampler` (lightworks/emulator/backends/fock_backend.py:67-87) hands `SamplerRunner`
(emulator/simulation/sampler.py:45-349) a `dict[State, float]` of every output state and the runner loops over it in
Python.  After K1/K3 that dict - one `State` object and one string hash per output - is >99 % of the wall time of a
Sampler task and makes BASELINE config 3 (92.6 M outputs) impossible.  This module runs the same steps on arrays:

    distribution_calculator (:76-106)    -> `distribution_arrays`: (keys uint8[K, m], vals[K]) in the reference's dict order
    _sample_N_outputs       (:240-327)   -> `sample_n_outputs`: threshold detection, herald check / removal, min_detection,
                                            post-selection rules and the merge of equal states as vectorised numpy, then the
                                            very `rng.choice` call of `_gen_samples_from_dist` (:329-349): same seed, same samples
    _sample_N_inputs        (:144-238)   -> `sample_n_inputs`: the same draw; detector, heralds and post-selection applied per
                                            drawn state (the reference's own `Detector._get_output` when it is not ideal, in
                                            the reference's order, so its `random` stream is consumed identically)

Two regimes.  Up to REGIME_A_MAX basis states the arrays live on the host and every step reproduces the reference bit for
bit (order of the dict, order of the floating-point additions, random streams).  Beyond it (perfect source, lossless
circuit) nothing of the distribution ever reaches the host: `lwb_basis_sample[_selected]` computes it, applies the
selection as a device mask over the whole basis, builds the CDF and draws on the GPU; only the drawn states are turned
into `State` objects.  Sampling the masked FULL-state distribution and mapping each drawn state (threshold detection,
herald removal) has the same law as the reference's draw from the merged dict, but not the same random stream.

Only `State` objects for DRAWN outcomes are ever created; `task.probability_distribution` is a lazy mapping.
There is no CPU fallback for the numerics: every probability comes from liblwb200.so.
"""
from __future__ import annotations

from collections.abc import Sequence

import numpy as np

from . import _lib

REGIME_A_MAX = 1 << 22  # basis states up to which the distribution is handled as host arrays (exact reference semantics)


class ArrayDistribution:
    """The output distribution of a Sampler as arrays.  `keys is None`: implicit full basis of (m, n) on the device
    (regime B) described by `basis = (backend_id, U_full, input_state)`."""

    def __init__(self, keys, vals, n_modes, basis=None):
        self.keys, self.vals, self.n_modes, self.basis = keys, vals, n_modes, basis

    def __len__(self):
        return len(self.vals) if self.vals is not None else -1


def make_lazy_distribution_class(base, state_cls):
    """A `ProbabilityDistribution` (sdk/results/probability_distribution.py:19-32: a dict subclass that forbids
    assignment) whose entries are only materialised when somebody looks at them."""

    class LazyProbabilityDistribution(base):
        def __init__(self, dist: ArrayDistribution, backend=None):
            dict.__init__(self)
            self._dist, self._backend, self._ready = dist, backend, False

        @property
        def arrays(self):
            """(keys uint8[K, m], vals float64[K]) in the reference's dict order, without building State objects."""
            d = self._dist
            if d.keys is None:  # regime B: compute and filter now
                backend_id, U, ins, threshold = d.basis  # noqa: N806
                ctx = self._backend.ctx
                p = ctx.perm_basis_probs(U, ins) if backend_id == _lib.BACKEND_PERMANENT else ctx.slos_basis_probs(U, ins)
                idx = np.nonzero(p > threshold)[0]
                off = _lib.keyspace_size(d.n_modes, int(sum(ins)) - 1) if sum(ins) else 0
                if backend_id == _lib.BACKEND_PERMANENT:
                    keys = _lib.keyspace_states(d.n_modes, int(sum(ins)), idx + off)
                else:
                    keys = _lib.slos_states(d.n_modes, int(sum(ins)), idx)
                d.keys, d.vals = keys, p[idx]
            return d.keys, d.vals

        def _fill(self):
            if not self._ready:
                keys, vals = self.arrays
                dict.update(self, {state_cls(k): v for k, v in zip(keys.tolist(), vals.tolist())})
                self._ready = True

        def __len__(self):
            if self._ready or self._dist.keys is None:
                self._fill()
                return dict.__len__(self)
            return len(self._dist.vals)

        def __iter__(self):
            self._fill()
            return dict.__iter__(self)

        def __getitem__(self, k):
            self._fill()
            return dict.__getitem__(self, k)

        def __contains__(self, k):
            self._fill()
            return dict.__contains__(self, k)

        def keys(self):
            self._fill()
            return dict.keys(self)

        def values(self):
            self._fill()
            return dict.values(self)

        def items(self):
            self._fill()
            return dict.items(self)

        def get(self, k, default=None):
            self._fill()
            return dict.get(self, k, default)

        def __eq__(self, other):
            self._fill()
            if hasattr(other, "_fill"):
                other._fill()
            return dict.__eq__(self, other)

        def __ne__(self, other):  # dict defines its own __ne__: without this, != would compare the unfilled dicts
            eq = self.__eq__(other)
            return eq if eq is NotImplemented else not eq

        __hash__ = None

        def __repr__(self):
            if not self._ready:
                n = "implicit full basis" if self._dist.keys is None else f"{len(self._dist.vals)} states"
                return f"lightworks.ProbabilityDistribution(<lazy, {n}>)"
            return base.__repr__(self)

    return LazyProbabilityDistribution


class LazyStates(Sequence):
    """A list of `State` objects over an occupation array uint8[K, m]: objects are created when an element is asked
    for.  What `SimulationResult.outputs` / `SamplingResult.outputs` hold when a task had thousands of outputs."""

    def __init__(self, rows: np.ndarray, state_cls):
        self.rows, self._cls, self._cache = rows, state_cls, None

    def _all(self):
        if self._cache is None:
            self._cache = [self._cls(r) for r in self.rows.tolist()]
        return self._cache

    def __len__(self):
        return len(self.rows)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return self._all()[i]
        if self._cache is not None:
            return self._cache[i]
        return self._cls(self.rows[i].tolist())

    def __iter__(self):
        return iter(self._all())

    def __eq__(self, other):
        if isinstance(other, LazyStates):
            return np.array_equal(self.rows, other.rows)
        return self._all() == other

    __hash__ = None

    def __repr__(self):
        return repr(self._all())


def make_lazy_sampling_result_class(base, state_cls):
    """A `SamplingResult` (sdk/results/sampling_result.py:33-75) over (rows uint8[K, m], counts[K]): the {State: count}
    dict - one string hash per distinct outcome - is built when dict access is used."""

    class LazySamplingResult(base):
        def __init__(self, rows, counts, input):  # noqa: A002
            if not isinstance(input, state_cls):
                base.__init__(self, {}, input)  # raises the reference's error
            dict.__init__(self)
            self.rows, self.counts, self._ready = rows, counts, False
            setattr(self, f"_{base.__name__}__input", input)
            setattr(self, f"_{base.__name__}__outputs", LazyStates(rows, state_cls))

        def _fill(self):
            if not self._ready:
                self._ready = True
                for st, c in zip(self.outputs, self.counts.tolist()):
                    dict.__setitem__(self, st, int(c))

        def __len__(self):
            return len(self.counts)

        def __iter__(self):
            self._fill()
            return dict.__iter__(self)

        def __contains__(self, k):
            self._fill()
            return dict.__contains__(self, k)

        def __getitem__(self, item):
            self._fill()
            return base.__getitem__(self, item)

        def keys(self):
            self._fill()
            return dict.keys(self)

        def values(self):
            if not self._ready:
                return [int(c) for c in self.counts.tolist()]
            return dict.values(self)

        def items(self):
            self._fill()
            return dict.items(self)

        def get(self, k, default=None):
            self._fill()
            return dict.get(self, k, default)

        def __eq__(self, other):
            self._fill()
            if hasattr(other, "_fill"):
                other._fill()
            return dict.__eq__(self, other)

        def __ne__(self, other):  # dict defines its own __ne__: without this, != would compare the unfilled dicts
            eq = self.__eq__(other)
            return eq if eq is NotImplemented else not eq

        __hash__ = None

        def __repr__(self):
            self._fill()
            return base.__repr__(self)

    return LazySamplingResult


def make_lazy_simulation_result_class(base):
    """A `SimulationResult` (sdk/results/simulation_result.py:33-90) that keeps the amplitude / probability array and
    the state lists but only builds the nested dict[State][State] - one string hash per output state, >95 % of the
    wall time of a Simulator / Analyzer task after the batched kernel - when dict access is actually used.
    `.array`, `.inputs`, `.outputs`, `.result_type` and everything defined on the base class work unchanged."""

    class LazySimulationResult(base):
        def __init__(self, results, result_type, inputs, outputs, **kwargs):
            if result_type not in {"probability", "probability_amplitude"}:
                base.__init__(self, results, result_type, inputs, outputs)  # raises the reference's error
            array = np.array(results)
            if len(inputs) != array.shape[0] or len(outputs) != array.shape[1]:
                base.__init__(self, results, result_type, inputs, outputs)  # raises the reference's error
            dict.__init__(self)
            # the private attributes of the base class (name-mangled there)
            for name, value in (("result_type", result_type), ("array", array), ("inputs", inputs), ("outputs", outputs)):
                setattr(self, f"_{base.__name__}__{name}", value)
            self._ready = False
            for k, v in kwargs.items():
                setattr(self, k, v)

        def _fill(self):
            if not self._ready:
                self._ready = True
                arr, outs = self.array, self.outputs
                for i, istate in enumerate(self.inputs):
                    dict.__setitem__(self, istate, dict(zip(outs, arr[i])))

        def __len__(self):
            return len(self.inputs) if not self._ready else dict.__len__(self)

        def __iter__(self):
            self._fill()
            return dict.__iter__(self)

        def __contains__(self, k):
            self._fill()
            return dict.__contains__(self, k)

        def __getitem__(self, item):
            self._fill()
            return base.__getitem__(self, item)

        def keys(self):
            self._fill()
            return dict.keys(self)

        def values(self):
            self._fill()
            return dict.values(self)

        def items(self):
            self._fill()
            return dict.items(self)

        def get(self, k, default=None):
            self._fill()
            return dict.get(self, k, default)

        def __eq__(self, other):
            self._fill()
            if hasattr(other, "_fill"):
                other._fill()
            return dict.__eq__(self, other)

        def __ne__(self, other):  # dict defines its own __ne__: without this, != would compare the unfilled dicts
            eq = self.__eq__(other)
            return eq if eq is NotImplemented else not eq

        __hash__ = None

        def __repr__(self):
            self._fill()
            return base.__repr__(self)

    return LazySimulationResult


# ---------------------------------------------------------------------------------------------------------------
def _merge_first_seen(rows: np.ndarray, weights: np.ndarray):
    """dict accumulation `d[k] = d.get(k, 0) + w` over rows in order: unique rows in first-seen order and their sums,
    added sequentially in input order (np.bincount accumulates in index order)."""
    if len(rows) == 0:
        return rows, weights
    rows = np.ascontiguousarray(rows)
    if rows.dtype != np.uint8 and rows.size and rows.max() < 256 and rows.min() >= 0:
        rows = rows.astype(np.uint8)
    # equal rows are found by comparing them as byte strings (one memcmp per comparison, not a lexicographic row sort)
    as_bytes = rows.view(np.dtype((np.void, rows.dtype.itemsize * rows.shape[1]))).reshape(-1)
    _, first, inv = np.unique(as_bytes, return_index=True, return_inverse=True)
    inv = inv.reshape(-1)
    order = np.argsort(first, kind="stable")
    pos = np.empty(len(first), dtype=np.int64)
    pos[order] = np.arange(len(first))
    sums = np.bincount(pos[inv], weights=weights, minlength=len(first))
    return rows[first[order]], sums


def distribution_arrays(backend, circuit, inputs: dict, threshold: float):
    """pdist_calc for State inputs (simulation/probability_distribution.py:25-73) on arrays; `inputs` is what
    `Source._build_statistics` returned (perfect source or brightness only).  Returns ArrayDistribution."""
    n_modes = circuit.n_modes
    items = [(list(k.s) if hasattr(k, "s") else list(k), float(p)) for k, p in inputs.items()]
    if len(items) == 1 and items[0][1] == 1 and circuit.loss_modes == 0:
        ins = items[0][0]
        n = sum(ins)
        if n > 0 and _lib.fock_count(n_modes, n) > REGIME_A_MAX:
            return ArrayDistribution(None, None, n_modes, basis=(backend._backend_id, circuit.U_full, ins, threshold))
    keys_all, vals_all = [], []
    for ins, prob in items:
        keys, vals = backend.full_probability_arrays(circuit, ins)
        if not keys_all and prob == 1:
            keys_all.append(keys)
            vals_all.append(vals)          # `pdist = sub_dist` (:57-58)
        else:
            keys_all.append(keys)
            vals_all.append(vals * prob)   # `p * prob` (:60, :64-66)
    if len(keys_all) == 1:
        keys, vals = keys_all[0], vals_all[0]
    else:
        keys, vals = _merge_first_seen(np.concatenate(keys_all), np.concatenate(vals_all))
    if circuit.loss_modes > 0:
        # `sum(pdist.values())` (:70).  The permanent backend's values are Python floats, which CPython >= 3.12 adds
        # with Neumaier compensation; the SLOS backend's are numpy scalars, added plainly left to right.
        total = sum(vals.tolist()) if backend._backend_id == _lib.BACKEND_PERMANENT else float(np.cumsum(vals)[-1])
        if total < 1:               # :71-73: assigns (and so may overwrite) the vacuum entry
            vac = np.nonzero(~keys.any(axis=1))[0]
            if len(vac):
                vals = vals.copy()
                vals[vac[0]] = 1 - total
            else:
                keys = np.concatenate([keys, np.zeros((1, n_modes), dtype=np.uint8)])
                vals = np.concatenate([vals, [1 - total]])
    if len(vals) == 0:  # sampler.py:99-100
        keys, vals = np.zeros((1, n_modes), dtype=np.uint8), np.ones(1)
    return ArrayDistribution(np.ascontiguousarray(keys), np.ascontiguousarray(vals, dtype=np.float64), n_modes)


class Selection:
    """The selection criteria of one Sampler task in array form (heralds on the output, min_detection, post-selection
    rules mapped to full-circuit modes); `rules is None` when the post-selection is an arbitrary Python function."""

    def __init__(self, n_modes, heralds: dict, min_detection: int, post_selection, photon_counting: bool, state_cls=list):
        self.n_modes, self.heralds, self.min_detection = n_modes, dict(heralds), int(min_detection)
        self.post_selection, self.photon_counting, self.state_cls = post_selection, bool(photon_counting), state_cls
        self.kept = [j for j in range(n_modes) if j not in self.heralds]  # modes left after herald removal
        rules = getattr(post_selection, "rules", None)
        if rules is not None:
            self.rules = [(tuple(self.kept[m] for m in r.modes), tuple(int(c) for c in r.n_photons)) for r in rules]
        elif type(post_selection).__name__ == "DefaultPostSelection":
            self.rules = []
        else:
            self.rules = None

    def mask(self, states: np.ndarray) -> np.ndarray:
        """Which (already threshold-detected) full-width states pass heralds, min_detection and post-selection."""
        ok = np.ones(len(states), dtype=bool)
        for m, n in self.heralds.items():
            ok &= states[:, m] == n
        reduced = states[:, self.kept]
        ok &= reduced.sum(axis=1) >= self.min_detection
        if self.rules is not None:
            for modes, counts in self.rules:
                ok &= np.isin(states[:, list(modes)].sum(axis=1), counts)
        else:  # arbitrary function: one Python call per candidate state, on the herald-free state like the reference
            idx = np.nonzero(ok)[0]
            keep = np.fromiter((bool(self.post_selection.validate(self.state_cls(s))) for s in reduced[idx].tolist()),
                               dtype=bool, count=len(idx))
            ok[idx] = keep
        return ok


def _draw(probs: np.ndarray, n_samples: int, seed, normalize: bool, process_seed):
    """_gen_samples_from_dist (:329-349) on an array of probabilities: drawn indices."""
    probs = np.array(probs, dtype=float)
    if normalize:
        probs /= np.cumsum(probs)[-1]  # `probs /= sum(probs)`: the builtin sum adds numpy scalars left to right
    rng = np.random.default_rng(process_seed(seed))
    return rng.choice(len(probs), p=probs, size=n_samples)


def _counts_first_seen(samples: np.ndarray):
    """Counter(samples): distinct values in first-seen order and their counts."""
    uniq, first, cnt = np.unique(samples, return_index=True, return_counts=True)
    order = np.argsort(first, kind="stable")
    return uniq[order].astype(np.int64), cnt[order].astype(np.int64)


def sample_n_outputs(dist: ArrayDistribution, sel: Selection, n_samples: int, seed, process_seed, no_states_error):
    """_sample_N_outputs (:240-327), regime A: (rows, counts) of the drawn states in the reference's order."""
    keys = dist.keys if sel.photon_counting else np.minimum(dist.keys, 1)
    ok = sel.mask(keys)
    rows, probs = keys[ok][:, sel.kept], dist.vals[ok]
    if sel.heralds or not sel.photon_counting:  # only then can two kept states coincide (:311-314)
        rows, probs = _merge_first_seen(rows, probs)
    if len(rows) == 0:
        raise no_states_error
    idx, cnt = _counts_first_seen(_draw(probs, n_samples, seed, True, process_seed))
    return rows[idx], cnt


def sample_n_inputs(dist: ArrayDistribution, sel: Selection, n_samples: int, seed, process_seed, state_cls, detector):
    """_sample_N_inputs (:144-238), regime A.  Returns ({State: count}, renormalised vals or None)."""
    renorm = None
    try:
        samples = _draw(dist.vals, n_samples, seed, False, process_seed)
    except ValueError as e:
        total_p = sum(dist.vals.tolist())
        if abs(total_p - 1) > 0.01:
            msg = f"Probability distribution significantly deviated from required normalisation ({total_p})."
            raise ValueError(msg) from e
        samples = _draw(dist.vals, n_samples, seed, True, process_seed)
        renorm = dist.vals / total_p
    idx, cnt = _counts_first_seen(samples)
    rows, counts = _detect_and_select(dist.keys[idx], cnt, sel, seed, state_cls, detector)
    return rows, counts, renorm


def _detect_and_select(states: np.ndarray, counts: np.ndarray, sel: Selection, seed, state_cls, detector):
    """The per-sample tail of _sample_N_inputs (:206-238): detector, herald check and removal, post-selection,
    min_detection, counting.  `states` are the distinct drawn full-width states in first-drawn order."""
    detector._set_random_seed(seed)
    ideal = detector.efficiency == 1 and detector.p_dark == 0 and detector.photon_counting
    if not ideal:
        # the reference's own detector, one call per sample in the reference's order: same `random` stream
        out = [detector._get_output(state_cls(s)).s for s, c in zip(states.tolist(), counts.tolist()) for _ in range(c)]
        states = np.array(out, dtype=np.int64).reshape(len(out), states.shape[1])
        counts = np.ones(len(out), dtype=np.int64)
    ok = sel.mask(states)
    rows, sums = states[ok][:, sel.kept], counts[ok]
    if sel.heralds or not ideal:  # herald removal / detector noise can map distinct drawn states onto one
        rows, sums = _merge_first_seen(rows, sums.astype(np.float64))
    return np.ascontiguousarray(rows, dtype=np.uint8), np.asarray(sums).astype(np.int64)


# ---------------------------------------------------------------------------------------------------------------
def sample_device(backend, dist: ArrayDistribution, sel: Selection, n_samples: int, seed, process_seed, state_cls, detector,
                  mode: str, no_states_error):
    """Regime B: the distribution stays on the GPU.  mode "output": draws from the states that pass the selection
    (device mask); mode "input": draws from the whole distribution, selection applied to the drawn states."""
    backend_id, U, ins, threshold = dist.basis  # noqa: N806
    ctx = backend.ctx
    m, n = dist.n_modes, int(sum(ins))
    rng = np.random.default_rng(process_seed(seed))
    uniform = rng.random(int(n_samples))
    if mode == "output":
        if sel.rules is None:
            raise ValueError("post-selection by an arbitrary function is not available for distributions kept on the "
                             f"device (more than {REGIME_A_MAX} basis states); use PostSelection rules")
        idx, total = ctx.basis_sample_selected(U, ins, uniform, threshold, backend_id, not sel.photon_counting,
                                               sel.min_detection, sel.heralds, sel.rules)
        if not total > 0:
            raise no_states_error
    else:
        idx, total = ctx.basis_sample(U, ins, uniform, threshold, backend_id)
        if abs(total - 1) > 0.01:
            raise ValueError(f"Probability distribution significantly deviated from required normalisation ({total}).")
    uidx, cnt = _counts_first_seen(idx.astype(np.int64))
    if backend_id == _lib.BACKEND_PERMANENT:
        off = _lib.keyspace_size(m, n - 1) if n else 0
        states = _lib.keyspace_states(m, n, uidx.astype(np.uint64) + off)
    else:
        states = _lib.slos_states(m, n, uidx)
    if mode == "output":
        if not sel.photon_counting:
            states = np.minimum(states, 1)
        rows, sums = states[:, sel.kept], cnt
        if sel.heralds or not sel.photon_counting:
            rows, sums = _merge_first_seen(rows, cnt.astype(np.float64))
        return np.ascontiguousarray(rows, dtype=np.uint8), np.asarray(sums).astype(np.int64)
    return _detect_and_select(states.astype(np.int64), cnt, sel, seed, state_cls, detector)

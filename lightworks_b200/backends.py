"""
Drop-in backends for the Lightworks emulator running on liblwb200.so (sm_100a).

Mirrors the operator interface of the reference (same names, arguments and error behaviour):
  lightworks/emulator/backends/fock_backend.py:43-124   FockBackend (the 4 overridable methods)
  lightworks/emulator/backends/permanent.py:30-163      PermanentBackend
  lightworks/emulator/backends/slos.py:28-105           SLOSBackend
  lightworks/emulator/backends/backend.py:27-122        Backend (name -> class dispatch)

Every numerical result comes from the CUDA library; there is no CPU fallback.  The classes work
stand-alone (lists / numpy in, dict[State, float] out, with lightworks_b200.state.State keys) and,
after `install()`, as subclasses of the reference's FockBackend inside an unmodified Lightworks SDK.
"""
from __future__ import annotations

from collections.abc import Mapping

import numpy as np

from . import _lib
from .state import State

_DTYPES = {"complex128": _lib.LWB_C128, "complex64": _lib.LWB_C64, _lib.LWB_C128: _lib.LWB_C128, _lib.LWB_C64: _lib.LWB_C64}

# Rebound to lightworks.State / settings by install(); module-level so the mixins stay import-light.
_state_cls = State
_settings = None
_convolve_on_gpu = False
_native_source = True
_array_sampler = True  # Sampler tasks run the array-native runner (array_sampler.py); False: the reference's runner
_n_gpus = 1          # install(n_gpus=N) / set_n_gpus(N): GPUs one process spreads the hot path over
_comm = None         # the LOCAL communicator shared by every backend object of this process


def set_n_gpus(n_gpus: int) -> None:
    """Spread the hot path of every B200 backend of this process over `n_gpus` GPUs: the output
    states of PermanentBackend.full_probability_distribution (permanent.py:145-157) are sharded by equal-work rank
    ranges, the unique sub-inputs of an imperfect source (probability_distribution.py:127-133) run round-robin.
    1 restores the single-device path."""
    global _n_gpus, _comm
    n_gpus = int(n_gpus)
    if n_gpus < 1:
        raise ValueError("n_gpus must be >= 1")
    if n_gpus != _n_gpus and _comm is not None:
        _comm.close()
        _comm = None
    _n_gpus = n_gpus


def local_comm():
    """The process-wide LOCAL communicator (None on one GPU); created on first use, fails loudly without the devices."""
    global _comm
    if _n_gpus > 1 and _comm is None:
        _comm = _lib.Comm.local(_n_gpus)
    return _comm


def set_native_source_statistics(flag: bool) -> None:
    """After install(), `Source._build_statistics` (emulator/components/source.py:134-161) is answered by the native
    enumeration `lwb_source_statistics` (same inputs, same dict order, probabilities equal to ~1e-16; 1.4 s instead of
    131 s for 8 photons with all three imperfections).  False restores the reference's Python loops."""
    global _native_source
    _native_source = bool(flag)


def set_array_sampler(flag: bool) -> None:
    """After install(), Sampler tasks on a B200 backend run `lightworks_b200.array_sampler` (same distribution, same
    seeded samples, no `State` object per output state) instead of the reference's SamplerRunner.  On by default;
    False restores the reference's runner fed by `full_probability_distribution`."""
    global _array_sampler
    _array_sampler = bool(flag)


def set_convolve_on_gpu(flag: bool) -> None:
    """Opt in to combining the source statistics on the device: after install(), a Sampler on a B200 backend then
    routes `pdist_calc` (simulation/probability_distribution.py:25-164, called at simulation/sampler.py:96) through
    lightworks_b200.convolution.pdist_calc.  Off by default because the returned dict is in key-space order, not
    in the reference's first-merged order (same distribution, different seeded samples)."""
    global _convolve_on_gpu
    _convolve_on_gpu = bool(flag)


def _threshold() -> float:
    """settings.sampler_probability_threshold, read at call time (permanent.py:151, slos.py:74)."""
    if _settings is not None:
        return float(_settings.sampler_probability_threshold)
    return 1e-9


def _as_list(state) -> list[int]:
    # the reference accepts State or list for both states (backend_test.py:145-151,172-178)
    return list(state.s) if hasattr(state, "s") else [int(v) for v in state]


class LazyDistribution(Mapping):
    """dict[State, float] view over (keys uint8[K, m], vals[K]) that only builds State objects on
    demand; iteration order = the reference's dict insertion order.  `.keys_array` / `.vals_array`
    give the array-native form."""

    def __init__(self, keys: np.ndarray, vals: np.ndarray, state_cls=None):
        self.keys_array, self.vals_array = keys, vals
        self._cls = state_cls or _state_cls
        self._index = None

    def _build(self):
        if self._index is None:
            self._index = {self._cls(k): i for i, k in enumerate(self.keys_array.tolist())}
        return self._index

    def __len__(self):
        return len(self.vals_array)

    def __iter__(self):
        return iter(self._build())

    def __getitem__(self, key):
        return float(self.vals_array[self._build()[key]])

    def to_dict(self) -> dict:
        return {self._cls(k): float(v) for k, v in zip(self.keys_array.tolist(), self.vals_array)}


class _B200Base:
    """State shared by both backends: the device context and the arithmetic dtype."""

    def __init__(self, device: int | None = None, dtype: str | int = "complex128"):
        self._device = device
        self._dtype = _DTYPES[dtype]
        self._ctx = None

    @property
    def ctx(self) -> _lib.Context:
        if self._ctx is None or not self._ctx.handle:  # a context borrowed from a closed communicator is dead: resolve again
            comm = local_comm() if self._device is None else None
            # with a LOCAL communicator device 0's context is the communicator's own (one set of buffers per device)
            self._ctx = comm.context(0) if comm is not None else _lib.default_context(self._device)
        return self._ctx

    def _pdist_arrays(self, circuit, input_state, backend_id):
        U = circuit.U_full  # noqa: N806
        ins = _as_list(input_state)
        if len(ins) != circuit.n_modes:
            raise ValueError("input state length does not match the circuit")
        comm = local_comm() if self._device is None else None
        if comm is not None and backend_id == _lib.BACKEND_PERMANENT and sum(ins) >= 7:
            # permanents sharded over the GPUs (SLOS layers depend on the whole previous layer: replicas only)
            return comm.pdist(U, circuit.n_modes, ins, _threshold(), self._dtype)
        return self.ctx.pdist(U, circuit.n_modes, ins, _threshold(), backend_id, self._dtype)


    _backend_id = _lib.BACKEND_PERMANENT

    def sample_outputs(self, circuit, input_state, n_samples: int, seed=None) -> dict:
        """On-device replacement of `Sampler(circuit, input_state, n_samples, random_seed=seed)` for the plain case
        (perfect source, photon-counting ideal detector, no heralds / loss / post-selection): the chain
        full_probability_distribution -> SamplerRunner._sample_N_outputs -> _gen_samples_from_dist
        (simulation/sampler.py:240-349) with the distribution computed, thresholded, normalised and searched on the
        GPU.  Only the n_samples uniform variates go in and the drawn positions come back; the variates come from
        numpy's `default_rng(seed).random(n_samples)`, the stream `rng.choice` consumes, so a seed draws the same
        states as the reference.  Returns {State: count} in first-drawn order (Counter semantics)."""
        from collections import Counter

        if getattr(circuit, "loss_modes", 0) or len(_as_list(input_state)) != circuit.n_modes:
            raise ValueError("sample_outputs covers lossless circuits with a full-width input state")
        ins = _as_list(input_state)
        rng = np.random.default_rng(seed)
        uniform = rng.random(int(n_samples))
        idx, _total = self.ctx.basis_sample(circuit.U_full, ins, uniform, _threshold(), self._backend_id)
        m, n = circuit.n_modes, sum(ins)
        unrank = _lib.fock_unrank if self._backend_id == _lib.BACKEND_PERMANENT else _lib.slos_unrank
        return {_state_cls(unrank(m, n, int(r)).tolist()): c for r, c in Counter(idx.tolist()).items()}


class PermanentBackendMixin(_B200Base):
    """PermanentBackend on the GPU: batched Glynn/Gray-code permanents (permanent.py:30-163)."""

    @property
    def name(self) -> str:
        return "permanent"

    @property
    def compatible_tasks(self) -> tuple[str, ...]:
        return ("Sampler", "Analyzer", "Simulator")

    def state_generator(self, n_modes: int, n_photons: int) -> list[list[int]]:
        """fock_basis(n_modes, n_photons), enumerated on the device (emulator/utils/state.py:22-35)."""
        if n_photons == 0:
            return [[0] * n_modes]
        return self.ctx.fock_basis(n_modes, n_photons).tolist()

    def state_generator_array(self, n_modes: int, n_photons: int) -> np.ndarray:
        """fock_basis as a read-only uint8 array [S, m]; the last few are kept (a Batch asks for the same one per task)."""
        if n_photons == 0:
            return np.zeros((1, n_modes), dtype=np.uint8)
        cache = self.__dict__.setdefault("_basis_cache", {})
        arr = cache.get((n_modes, n_photons))
        if arr is None:
            arr = self.ctx.fock_basis(n_modes, n_photons)
            arr.setflags(write=False)
            if arr.nbytes <= 64 << 20:
                if len(cache) >= 8:
                    cache.pop(next(iter(cache)))
                cache[(n_modes, n_photons)] = arr
        return arr

    def probability_amplitude(self, unitary, input_state, output_state) -> complex:
        """permanent.py:53-83"""
        return complex(self.amplitude_matrix(unitary, [_as_list(input_state)], [_as_list(output_state)])[0, 0])

    def probability(self, unitary, input_state, output_state) -> float:
        """permanent.py:85-114: abs(amplitude) ** 2"""
        return abs(self.probability_amplitude(unitary, input_state, output_state)) ** 2

    @staticmethod
    def _state_array(states) -> np.ndarray:
        if isinstance(states, np.ndarray):
            return np.ascontiguousarray(states, dtype=np.uint8)
        return np.array([_as_list(s) for s in states], dtype=np.uint8)

    def amplitude_matrix(self, unitary, inputs, outputs) -> np.ndarray:
        """All <out|U|in> at once: the double loop of simulation/simulator.py:98-104.  States as lists of State /
        list, or as one uint8 array [K, m]."""
        return self.ctx.perm_amplitudes(unitary, self._state_array(inputs), self._state_array(outputs), self._dtype, probs=False)

    def probability_matrix(self, unitary, inputs, outputs) -> np.ndarray:
        """All |<out|U|in>|^2 at once: the no-loss branch of simulation/analyzer.py:120-125."""
        return self.ctx.perm_amplitudes(unitary, self._state_array(inputs), self._state_array(outputs), self._dtype, probs=True)

    def full_probability_arrays(self, circuit, input_state):
        """(keys, vals) of full_probability_distribution without building Python State objects."""
        return self._pdist_arrays(circuit, input_state, _lib.BACKEND_PERMANENT)

    def full_probability_distribution(self, circuit, input_state) -> dict:
        """permanent.py:116-163"""
        keys, vals = self._pdist_arrays(circuit, input_state, _lib.BACKEND_PERMANENT)
        return {_state_cls(k): float(v) for k, v in zip(keys.tolist(), vals)}


class SLOSBackendMixin(_B200Base):
    """SLOSBackend on the GPU: layer-by-layer state-vector evolution (slos.py:28-138)."""

    _backend_id = _lib.BACKEND_SLOS

    @property
    def name(self) -> str:
        return "slos"

    @property
    def compatible_tasks(self) -> tuple[str, ...]:
        return ("Sampler",)

    def full_probability_arrays(self, circuit, input_state):
        return self._pdist_arrays(circuit, input_state, _lib.BACKEND_SLOS)

    def full_probability_distribution(self, circuit, input_state) -> dict:
        """slos.py:43-80"""
        keys, vals = self._pdist_arrays(circuit, input_state, _lib.BACKEND_SLOS)
        return {_state_cls(k): float(v) for k, v in zip(keys.tolist(), vals)}

    def calculate_arrays(self, unitary, input_state):
        """(states uint8[S, m], amplitudes complex128[S]) in the reference's dict insertion order."""
        ins = _as_list(input_state)
        U = np.ascontiguousarray(unitary, dtype=np.complex128)  # noqa: N806
        m, n = U.shape[0], sum(ins)
        amps = self.ctx.slos_amplitudes(U, ins, _lib.ORDER_SLOS)
        states = np.array([_lib.slos_unrank(m, n, r) for r in range(len(amps))], dtype=np.uint8).reshape(len(amps), m)
        return states, amps

    def calculate(self, unitary, input_state) -> dict:
        """slos.py:82-105: dict[tuple[int, ...], complex] in insertion order."""
        states, amps = self.calculate_arrays(unitary, input_state)
        return {tuple(k): complex(a) for k, a in zip(states.tolist(), amps)}


# ---- stand-alone classes (no lightworks needed) ---------------------------------------------------
class PermanentBackend(PermanentBackendMixin):
    pass


class SLOSBackend(SLOSBackendMixin):
    pass


class Backend:
    """Stand-alone mirror of lightworks.emulator.Backend's selection behaviour (backend.py:42-45,
    112-122) for the two hot-path methods; running Tasks needs the SDK, see install()."""

    _backend_map = (("permanent", PermanentBackend), ("slos", SLOSBackend))

    def __init__(self, backend: str) -> None:
        backends = dict(self._backend_map)
        if backend not in backends:
            msg = f"Backend name not recognised, valid options are: {', '.join(backends.keys())}."
            raise ValueError(msg)
        self._backend = backends[backend]()

    @property
    def backend(self) -> str:
        return self._backend.name

    def __str__(self) -> str:
        return self.backend


# ---- integration with an importable Lightworks SDK -------------------------------------------------
_installed = None


def install(dtype: str = "complex128", device: int | None = None, patch_names: bool = True, n_gpus: int | None = None):
    """Make `lightworks.emulator.Backend("permanent" | "slos")` run on the GPU.

    Rebinds Backend._backend_map (a plain class attribute, backend.py:42-45) to subclasses of the
    reference's FockBackend built here; the reference CPU backends stay reachable as
    "permanent_cpu" / "slos_cpu".  With patch_names, `lightworks.emulator.backends.PermanentBackend`
    / `.SLOSBackend` (the names user code imports) are rebound too; the reference classes stay in
    their defining modules.  n_gpus > 1: see set_n_gpus.  Returns (B200PermanentBackend, B200SLOSBackend).
    """
    global _installed, _state_cls, _settings
    import lightworks as lw
    from lightworks.emulator.backends.backend import Backend as LwBackend
    from lightworks.emulator.backends.fock_backend import FockBackend
    from lightworks.emulator.backends.permanent import PermanentBackend as RefPermanent
    from lightworks.emulator.backends.slos import SLOSBackend as RefSLOS
    from lightworks.emulator.simulation import AnalyzerRunner
    from lightworks.emulator.utils.exceptions import BackendError as LwBackendError
    from lightworks.emulator.utils.sim import check_photon_numbers
    from lightworks.sdk.results import SimulationResult
    from lightworks.sdk.tasks import Analyzer, Sampler, Simulator
    from lightworks.sdk.utils.exceptions import PhotonNumberError as LwPhotonNumberError
    from lightworks.sdk.utils.heralding import add_heralds_to_state
    from lightworks.sdk.utils.post_selection import DefaultPostSelection

    if n_gpus is not None:
        set_n_gpus(n_gpus)
    if _installed is not None:
        return _installed["classes"]
    _state_cls = lw.State
    _settings = lw.settings
    _lib._exc["BackendError"] = LwBackendError
    _lib._exc["PhotonNumberError"] = LwPhotonNumberError

    from . import array_sampler as asmp

    lazy_sim_cls = asmp.make_lazy_simulation_result_class(SimulationResult)

    class _BatchedAnalyzerRunner(AnalyzerRunner):
        """AnalyzerRunner with `_get_probs` (simulation/analyzer.py:111-153) evaluated as ONE batched
        device call per photon number instead of one Python call per (input, output) pair."""

        def __init__(self, data, backend):
            super().__init__(data, backend.probability, backend.state_generator)
            self._b = backend

        def run(self):
            """AnalyzerRunner.run (simulation/analyzer.py:60-108) returning the lazy SimulationResult: the steps and
            their order are the reference's; only the result object defers building dict[State][State]."""
            data = self.data
            inputs = list(data.inputs)
            check_photon_numbers(inputs, inputs[0].n_photons)
            pad = [0] * data.circuit.loss_modes
            full_inputs = [add_heralds_to_state(i, data.circuit.heralds.input) + pad for i in inputs]
            n_photons = sum(full_inputs[0]) - sum(data.circuit.heralds.output.values())
            full_outputs, filtered_outputs = self._generate_outputs(data.circuit.input_modes, n_photons)
            probs = self._get_probs(full_inputs, full_outputs)
            results = lazy_sim_cls(probs, "probability", inputs=inputs, outputs=filtered_outputs,
                                   performance=probs.sum() / len(full_inputs))
            if data.expected is not None:
                results.error_rate = self._calculate_error_rate(probs, inputs, filtered_outputs, data.expected)
            return results

        def _generate_outputs(self, n_modes, n_photons):
            """analyzer.py:189-225 on arrays when the post-selection is a set of rules (or none): outputs with and
            without the heralded modes; an arbitrary post-selection function keeps the reference's per-state loop."""
            data = self.data
            post = DefaultPostSelection() if data.post_selection is None else data.post_selection
            rules = getattr(post, "rules", None)
            if rules is None and type(post).__name__ != "DefaultPostSelection":
                return super()._generate_outputs(n_modes, n_photons)
            gen = self._b.state_generator_array
            if not data.circuit.loss_modes:
                arr = gen(n_modes, n_photons)
            else:
                arr = np.concatenate([gen(n_modes, n) for n in range(n_photons + 1)])
            ok = np.ones(len(arr), dtype=bool)
            for r in rules or []:
                ok &= np.isin(arr[:, list(r.modes)].sum(axis=1), list(r.n_photons))
            arr = arr[ok]
            if len(arr) == 0:
                raise ValueError("No valid outputs found, consider relaxing post-selection.")
            return _with_heralds(arr, data.circuit.heralds.output, 0), asmp.LazyStates(arr, lw.State)

        def _get_probs(self, full_inputs, full_outputs):
            circuit = self.data.circuit
            if circuit.loss_modes and isinstance(full_outputs, np.ndarray):
                full_outputs = full_outputs.tolist()
            U, loss = circuit.U_full, circuit.loss_modes  # noqa: N806
            probs = np.zeros((len(full_inputs), len(full_outputs)))
            if not loss:  # analyzer.py:121-125
                probs += self._b.probability_matrix(U, full_inputs, full_outputs)
                return probs
            n_in = sum(full_inputs[0])
            by_loss: dict[int, list[int]] = {}
            for j, outs in enumerate(full_outputs):
                n_loss = n_in - sum(outs)
                if n_loss < 0:  # analyzer.py:139-142
                    raise LwPhotonNumberError("Output photon number larger than input number.")
                by_loss.setdefault(n_loss, []).append(j)
            for n_loss, cols in by_loss.items():
                # every loss-mode completion of every output in this group (analyzer.py:144-151)
                loss_states = self.state_generator(loss, n_loss)
                expanded = [full_outputs[j] + ls for j in cols for ls in loss_states]
                p = self._b.probability_matrix(U, full_inputs, expanded)
                p = p.reshape(len(full_inputs), len(cols), len(loss_states))
                for c, j in enumerate(cols):
                    for q in range(len(loss_states)):  # same left-to-right accumulation as the reference
                        probs[:, j] += p[:, c, q]
            return probs

    def _with_heralds(states: np.ndarray, heralds: dict, pad: int) -> np.ndarray:
        """add_heralds_to_state (sdk/utils/heralding.py:20-52) + loss-mode padding for an array of states."""
        k, width = states.shape[0], states.shape[1] + len(heralds)
        out = np.zeros((k, width + pad), dtype=np.uint8)
        free = [j for j in range(width) if j not in heralds]
        out[:, free] = states
        for mode, count in heralds.items():
            out[:, mode] = count
        return out

    def _prepare_simulator(self, data):
        """The state bookkeeping of SimulatorRunner.run (simulator.py:70-100): (outputs, full inputs, full outputs)."""
        circuit = data.circuit
        in_heralds, out_heralds = circuit.heralds.input, circuit.heralds.output
        in_heralds_n, out_heralds_n = sum(in_heralds.values()), sum(out_heralds.values())
        target_n = data.inputs[0].n_photons + in_heralds_n
        check_photon_numbers(data.inputs, target_n - in_heralds_n)
        pad = [0] * circuit.loss_modes
        if data.outputs is None:  # every output state: stays an array, State objects are made on demand
            out_arr = self.state_generator_array(circuit.input_modes, target_n - out_heralds_n)
            outputs = asmp.LazyStates(out_arr, lw.State)
            full_outputs = _with_heralds(out_arr, out_heralds, circuit.loss_modes)
        else:
            check_photon_numbers(data.outputs, target_n - out_heralds_n)
            outputs = data.outputs
            full_outputs = np.array([add_heralds_to_state(o, out_heralds) + pad for o in outputs], dtype=np.uint8)
        full_inputs = np.array([add_heralds_to_state(i, in_heralds) + pad for i in data.inputs], dtype=np.uint8)
        return outputs, full_inputs, full_outputs

    def _run_simulator(self, task):
        """FockBackend.run_simulator + SimulatorRunner.run (fock_backend.py:53-58, simulator.py:58-111)
        with the inputs x outputs double loop replaced by one batched device call."""
        data = task._generate_task()
        outputs, full_inputs, full_outputs = _prepare_simulator(self, data)
        amplitudes = np.zeros((len(data.inputs), len(outputs)), dtype=complex)
        amplitudes += self.amplitude_matrix(data.circuit.U_full, full_inputs, full_outputs)
        return lazy_sim_cls(amplitudes, "probability_amplitude", inputs=data.inputs, outputs=outputs)

    def _run_simulator_batch(self, tasks):
        """lightworks.Batch of Simulator tasks (sdk/tasks/batch.py:45-58; run one by one at backends/backend.py:102-103):
        tasks whose circuits have the same size and whose full-width outputs coincide - a parameter scan, the
        tomography circuits - go to the device as ONE launch with a table of unitaries (lwb_perm_amplitudes_batch);
        anything else falls back to task-by-task.  Results in task order."""
        prepared = []
        for t in tasks:
            data = t._generate_task()
            prepared.append((data, *_prepare_simulator(self, data)))
        results = [None] * len(tasks)
        groups: dict[tuple, list[int]] = {}
        for i, (data, _outs, fin, fout) in enumerate(prepared):
            groups.setdefault((data.circuit.U_full.shape, fin.shape, fout.shape, fout.tobytes()), []).append(i)
        for idxs in groups.values():
            first = prepared[idxs[0]]
            U = np.stack([prepared[i][0].circuit.U_full for i in idxs])  # noqa: N806
            ins = np.stack([prepared[i][2] for i in idxs])
            amps = self.ctx.perm_amplitudes_batch(U, ins, first[3], self._dtype, probs=False)
            for k, i in enumerate(idxs):
                data, outs = prepared[i][0], prepared[i][1]
                results[i] = lazy_sim_cls(amps[k] + 0, "probability_amplitude", inputs=data.inputs, outputs=outs)
        return results

    def _run_analyzer(self, task):
        return _BatchedAnalyzerRunner(task._generate_task(), self).run()

    from lightworks.emulator.simulation.sampler import SamplerRunner
    from lightworks.sdk.results import ProbabilityDistribution, SamplingResult
    from lightworks.sdk.utils.exceptions import SamplerError
    from lightworks.sdk.utils.random import process_random_seed

    lazy_pd_cls = asmp.make_lazy_distribution_class(ProbabilityDistribution, lw.State)
    lazy_sr_cls = asmp.make_lazy_sampling_result_class(SamplingResult, lw.State)

    def _run_sampler(self, task):
        """FockBackend.run_sampler (fock_backend.py:67-87) + SamplerRunner (simulation/sampler.py:76-349) on arrays:
        same cache, same distribution and dict order, same seeded draws, but no State object per output state
        (lightworks_b200/array_sampler.py).  Annotated inputs (imperfect purity / indistinguishability) go through the
        device convolution when set_convolve_on_gpu(True), else through the reference's own runner."""
        data = task._generate_task()
        circuit = data.circuit
        runner = SamplerRunner(data, self.full_probability_distribution)  # resolves the default Source / Detector
        source, detector = runner.source, runner.detector
        cached = self._check_cache(data)
        if cached is not None and not isinstance(cached["pdist"], asmp.ArrayDistribution):
            return FockBackend.run_sampler(self, task)  # cached by the reference runner: let it reuse its dict
        if cached is not None:
            dist = cached["pdist"]
        else:
            if circuit.input_modes != len(data.input_state):  # sampler.py:82-85
                raise ValueError("Mismatch in number of modes between input and circuit.")
            input_state = lw.State(add_heralds_to_state(data.input_state, circuit.heralds.input))
            all_inputs = source._build_statistics(input_state)
            if all_inputs and isinstance(next(iter(all_inputs)), AnnotatedState):
                if not _convolve_on_gpu:
                    return FockBackend.run_sampler(self, task)
                from .convolution import pdist_calc as gpu_pdist_calc

                lazy = gpu_pdist_calc(circuit, all_inputs, self)
                keys, vals = lazy.keys_array, lazy.vals_array
                if len(vals) == 0:  # sampler.py:99-100
                    keys, vals = np.zeros((1, circuit.n_modes), dtype=np.uint8), np.ones(1)
                dist = asmp.ArrayDistribution(np.ascontiguousarray(keys), np.ascontiguousarray(vals), circuit.n_modes)
            else:
                dist = asmp.distribution_arrays(self, circuit, all_inputs, _threshold())
            self._add_to_cache(data, {"pdist": dist, "herald_cache": None})
        task._probability_distribution = lazy_pd_cls(dist, self)
        post_selection = DefaultPostSelection() if data.post_selection is None else data.post_selection
        min_detection = 0 if data.min_detection is None else data.min_detection
        heralds = circuit.heralds.output
        if heralds and max(heralds.values()) > 1 and not detector.photon_counting:  # sampler.py:208-216, :285-293
            raise SamplerError("Non photon number resolving detectors cannot be used when"
                               "a heralded mode has more than 1 photon.")
        output_mode = data.sampling_mode != "input"
        if output_mode and (detector.p_dark > 0 or detector.efficiency < 1):  # sampler.py:276-280
            raise SamplerError("To use detector dark counts or sub-unity detector efficiency "
                               "the sampling mode must be set to 'input'.")
        sel = asmp.Selection(circuit.n_modes, heralds, min_detection, post_selection,
                             detector.photon_counting if output_mode else True, lw.State)
        no_states = SamplerError("No output states compatible with provided post-selection/min-detection criteria.")
        if dist.keys is None:  # the distribution lives on the device
            rows, counts = asmp.sample_device(self, dist, sel, data.n_samples, data.random_seed, process_random_seed,
                                              lw.State, detector, "output" if output_mode else "input", no_states)
        elif output_mode:
            rows, counts = asmp.sample_n_outputs(dist, sel, data.n_samples, data.random_seed, process_random_seed, no_states)
        else:
            rows, counts, _ = asmp.sample_n_inputs(dist, sel, data.n_samples, data.random_seed, process_random_seed,
                                                   lw.State, detector)
        return lazy_sr_cls(rows, counts, data.input_state)

    def _run(self, task):
        # FockBackend.run is a multimethod object bound to the base implementations
        # (fock_backend.py:49-68), so dispatch by task type here.
        if isinstance(task, Simulator) and hasattr(self, "amplitude_matrix"):
            return _run_simulator(self, task)
        if isinstance(task, Analyzer) and hasattr(self, "probability_matrix"):
            return _run_analyzer(self, task)
        if isinstance(task, Sampler):
            if _array_sampler:
                return _run_sampler(self, task)  # array-native runner, shares the pdist cache (fock_backend.py:67-87)
            return FockBackend.run_sampler(self, task)
        raise LwBackendError("Task not supported on current backend.")

    class B200PermanentBackend(PermanentBackendMixin, FockBackend):
        def __init__(self):
            PermanentBackendMixin.__init__(self, device, dtype)

        run = _run
        run_simulator_batch = _run_simulator_batch

    from lightworks.sdk.tasks import Batch

    ref_run_batch = LwBackend._run_batch

    def _run_batch(self, task: Batch):
        """Backend._run_batch (backends/backend.py:101-103): a Batch made only of Simulator tasks on the B200
        permanent backend runs as one device launch; everything else keeps the reference's task-by-task loop."""
        tasks = list(task)
        be = self._backend
        if tasks and isinstance(be, B200PermanentBackend) and all(type(t) is Simulator for t in tasks):
            return be.run_simulator_batch(tasks)
        return [self._run(t) for t in tasks]

    # the dispatcher resolves annotations in module globals; Batch is local to install(), so hand it the class itself
    _run_batch.__annotations__ = {"task": Batch}
    LwBackend._run.register(_run_batch)

    class B200SLOSBackend(SLOSBackendMixin, FockBackend):
        def __init__(self):
            SLOSBackendMixin.__init__(self, device, dtype)

        run = _run

    import lightworks.emulator.backends as lw_backends_pkg
    import lightworks.emulator.simulation.sampler as lw_sampler_mod

    ref_pdist_calc = lw_sampler_mod.pdist_calc

    def _pdist_calc(circuit, inputs, probability_func):
        """simulation/sampler.py:96 calls this name; B200 backends may combine the inputs on the device."""
        owner = getattr(probability_func, "__self__", None)
        if _convolve_on_gpu and isinstance(owner, _B200Base):
            from .convolution import pdist_calc as gpu_pdist_calc

            return gpu_pdist_calc(circuit, inputs, owner).to_dict()
        return ref_pdist_calc(circuit, inputs, probability_func)

    lw_sampler_mod.pdist_calc = _pdist_calc

    from lightworks.emulator.components.source import Source as LwSource
    from lightworks.emulator.state import AnnotatedState

    ref_build_statistics = LwSource._build_statistics

    def _build_statistics(self, state):
        """components/source.py:134-161 through lwb_source_statistics (host only)."""
        if not _native_source:
            return ref_build_statistics(self, state)
        annotated, entries = _lib.source_statistics(list(state), self.purity, self.brightness, self.indistinguishability,
                                                    self.probability_threshold)
        if annotated:
            return {AnnotatedState([list(mode) for mode in k]): p for k, p in entries}
        return {lw.State(list(k)): p for k, p in entries}

    LwSource._build_statistics = _build_statistics

    _installed = {"old_map": LwBackend._backend_map, "classes": (B200PermanentBackend, B200SLOSBackend),
                  "Backend": LwBackend, "pkg": lw_backends_pkg, "patched": patch_names,
                  "ref": (RefPermanent, RefSLOS), "sampler_mod": lw_sampler_mod, "ref_pdist_calc": ref_pdist_calc,
                  "source_cls": LwSource, "ref_build_statistics": ref_build_statistics, "ref_run_batch": ref_run_batch}
    if patch_names:
        lw_backends_pkg.PermanentBackend = B200PermanentBackend
        lw_backends_pkg.SLOSBackend = B200SLOSBackend
    LwBackend._backend_map = (
        ("permanent", B200PermanentBackend),
        ("slos", B200SLOSBackend),
        ("permanent_cpu", RefPermanent),
        ("slos_cpu", RefSLOS),
    )
    return _installed["classes"]


def uninstall():
    """Restore the reference's own backends."""
    global _installed, _state_cls, _settings
    if _installed is None:
        return
    _installed["Backend"]._backend_map = _installed["old_map"]
    if _installed["patched"]:
        _installed["pkg"].PermanentBackend, _installed["pkg"].SLOSBackend = _installed["ref"]
    _installed["sampler_mod"].pdist_calc = _installed["ref_pdist_calc"]
    _installed["source_cls"]._build_statistics = _installed["ref_build_statistics"]
    _installed["Backend"]._run.register(_installed["ref_run_batch"])
    _installed = None
    _state_cls = State
    _settings = None
    _lib._exc["BackendError"] = _lib.BackendError
    _lib._exc["PhotonNumberError"] = _lib.PhotonNumberError

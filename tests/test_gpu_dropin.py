"""
Drop-in tests: the UNMODIFIED Lightworks SDK (tasks, circuits, source/detector models, result types)
running on the B200 backends through `lightworks_b200.install()`, compared with the reference's own
CPU backends in the same process.  Needs a GPU and an importable reference (baseline/_ref on the GPU
box, /root/reference in the build container); skipped otherwise.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.reference]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lw():
    from oracle import refshim

    lightworks = refshim.import_lightworks()
    import lightworks_b200

    lightworks_b200.install()
    yield lightworks
    lightworks_b200.uninstall()


def _close(a, b, rel=1e-10, abs_=1e-15):
    return np.all(np.abs(np.asarray(a) - np.asarray(b)) <= rel * np.abs(np.asarray(b)) + abs_)


def test_backend_selection(lw):
    from lightworks.emulator import Backend

    assert str(Backend("permanent")) == "permanent"
    assert type(Backend("permanent")._backend).__name__ == "B200PermanentBackend"
    assert type(Backend("slos")._backend).__name__ == "B200SLOSBackend"
    with pytest.raises(ValueError):
        Backend("not_a_backend")
    from lightworks.emulator import BackendError

    with pytest.raises(BackendError):  # slos is Sampler-only (backend_test.py:41-48)
        Backend("slos").run(lw.Simulator(lw.Unitary(lw.random_unitary(4)), lw.State([1, 0, 1, 0])))


@pytest.mark.parametrize("m,n,seed", [(6, 3, 1), (12, 6, 2)])
def test_simulator_c1_c2(lw, m, n, seed):
    from lightworks.emulator import Backend

    task = lw.Simulator(lw.Unitary(lw.random_unitary(m, seed=seed)), lw.State([1] * n + [0] * (m - n)))
    got, ref = Backend("permanent").run(task), Backend("permanent_cpu").run(task)
    assert got.outputs == ref.outputs and got.inputs == ref.inputs
    assert _close(got.array, ref.array)


def test_analyzer_c2_and_lossy_heralded(lw):
    from lightworks.emulator import Backend

    task = lw.Analyzer(lw.Unitary(lw.random_unitary(12, seed=2)), lw.State([1] * 6 + [0] * 6))
    got, ref = Backend("permanent").run(task), Backend("permanent_cpu").run(task)
    assert got.outputs == ref.outputs
    assert _close(got.array, ref.array)
    assert got.performance == pytest.approx(ref.performance, rel=1e-10)
    # lossy + herald + post-selection: loss-mode summation path (analyzer.py:136-151)
    circ = lw.PhotonicCircuit(5)
    circ.add(lw.Unitary(lw.random_unitary(5, seed=5)))
    for i in range(5):
        circ.loss(i, 0.15 + 0.05 * i)
    circ.herald(4, 0)
    an = lw.Analyzer(circ, [lw.State([1, 1, 0, 1]), lw.State([0, 1, 1, 1])])
    an.post_selection = lambda s: s[0] <= 1
    got, ref = Backend("permanent").run(an), Backend("permanent_cpu").run(an)
    assert got.outputs == ref.outputs
    assert _close(got.array, ref.array, 1e-10, 1e-16)
    assert got.performance == pytest.approx(ref.performance, rel=1e-10)


@pytest.mark.parametrize("name", ["permanent", "slos"])
def test_sampler_distribution_and_samples(lw, name):
    from lightworks.emulator import Backend, Detector, Source

    circ = lw.PhotonicCircuit(6)
    circ.add(lw.Unitary(lw.random_unitary(6, seed=8)))
    for i in range(6):
        circ.loss(i, 0.1)
    src = Source(purity=0.95, brightness=0.9, indistinguishability=0.9)
    mk = lambda: lw.Sampler(circ, lw.State([1, 0, 1, 0, 1, 0]), 2000, source=src,  # noqa: E731
                            detector=Detector(efficiency=0.9), sampling_mode="input", random_seed=7)
    t1, t2 = mk(), mk()
    r1 = Backend(name).run(t1)
    r2 = Backend(name + "_cpu").run(t2)
    d1, d2 = dict(t1.probability_distribution), dict(t2.probability_distribution)
    assert list(d1) == list(d2)  # same states in the same order
    vac = lw.State([0] * 6)
    for k in d2:
        if k == vac:
            assert d1[k] == pytest.approx(d2[k], abs=1e-12)
        else:
            assert d1[k] == pytest.approx(d2[k], rel=1e-10, abs=1e-16)
    # same seed, same ordered distribution -> the drawn samples agree unless a cumulative-sum edge moves
    c1, c2 = dict(r1), dict(r2)
    assert sum(c1.values()) == sum(c2.values())
    common = sum(min(c1.get(k, 0), c2.get(k, 0)) for k in set(c1) | set(c2))
    assert common >= 0.99 * sum(c2.values())


def test_sampler_cache_is_inherited(lw):
    # fock_backend.py:67-87 / backend_test.py:255-267
    from lightworks.emulator import Backend

    b = Backend("permanent")
    s = lw.Sampler(lw.Unitary(lw.random_unitary(4, seed=3)), lw.State([1, 0, 1, 0]), 100)
    b.run(s)
    assert b._backend._check_cache(s._generate_task()) is not None


def test_reference_emulator_suite_runs_on_b200():
    """The reference's own tests/emulator suite with Backend("permanent"/"slos"), PermanentBackend and
    SLOSBackend redirected to the B200 classes (SURVEY.md section 8d 'correctness gates')."""
    # the reference marks its statistical sampling tests `flaky` (rerun plugin, absent here): one retry of the suite
    for attempt in range(2):
        r = subprocess.run([sys.executable, "-m", "oracle.run_reference_tests", "--b200", "-x"], cwd=ROOT,
                           capture_output=True, text=True, timeout=1500)
        if r.returncode == 0:
            break
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail


def test_reference_qubit_and_tomography_suites_run_on_b200():
    """Indirect consumers of the hot path: the reference's qubit gate tests (Simulator on the
    permanent backend) and state / process tomography tests (Samplers on the slos backend), backends redirected."""
    r = subprocess.run([sys.executable, "-m", "oracle.run_reference_tests", "--b200", "qubit/gate_test.py",
                        "tomography/state_test.py", "tomography/process_test.py", "-k", "not plot", "-x"],
                       cwd=ROOT, capture_output=True, text=True, timeout=1500)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail


@pytest.mark.parametrize("name", ["permanent", "slos"])
def test_sample_outputs_equals_reference_sampler(lw, name):
    """On-device sampling (lwb_basis_sample) against the unmodified Sampler task on the reference CPU backend with
    the same seed: sampler.py:240-349 end to end."""
    from lightworks.emulator import Backend

    circ = lw.Unitary(lw.random_unitary(8, seed=5))
    st = lw.State([1, 0, 1, 0, 1, 0, 1, 0])
    task = lw.Sampler(circ, st, 5000, random_seed=21)
    ref = dict(Backend(name + "_cpu").run(task))
    got = Backend(name)._backend.sample_outputs(task._generate_task().circuit, st, 5000, seed=21)
    assert sum(got.values()) == 5000
    common = sum(min(got.get(k, 0), ref.get(k, 0)) for k in set(got) | set(ref))
    assert common >= 5000 - 2


@pytest.mark.parametrize("name", ["permanent", "slos"])
def test_imperfect_source_combined_on_device(lw, name):
    """Opt-in device path for pdist_calc (probability_distribution.py:25-164): an unmodified Sampler with an
    imperfect source on a lossy circuit, distribution compared entry by entry with the reference CPU backend."""
    import lightworks_b200
    from lightworks import emulator
    from lightworks.emulator import Backend

    circ = lw.PhotonicCircuit(5)
    for i, mode in enumerate([0, 2, 1, 3, 0, 2]):
        circ.bs(mode, reflectivity=0.3 + 0.1 * i, loss=0.05 * (i % 3))
        circ.ps(mode + 1, 0.2 * i)
    src = emulator.Source(purity=0.95, brightness=0.85, indistinguishability=0.9)
    mk = lambda: lw.Sampler(circ, lw.State([1, 0, 1, 1, 0]), 2000, source=src, random_seed=3)  # noqa: E731
    t1, t2 = mk(), mk()
    lightworks_b200.set_convolve_on_gpu(True)
    try:
        r1 = Backend(name).run(t1)
    finally:
        lightworks_b200.set_convolve_on_gpu(False)
    Backend(name + "_cpu").run(t2)
    d1, d2 = dict(t1.probability_distribution), dict(t2.probability_distribution)
    assert set(d1) == set(d2)
    vac = lw.State([0] * 5)
    for k, v in d2.items():
        assert d1[k] == pytest.approx(v, rel=1e-10, abs=1e-12 if k == vac else 1e-16)
    assert sum(dict(r1).values()) == 2000


def test_batch_of_simulators_runs_as_one_launch(lw):
    """lightworks.Batch across unitaries (sdk/tasks/batch.py:45-58, serial at backends/backend.py:102-103): a parameter
    scan of Simulator tasks goes to the device as ONE launch with a table of unitaries and equals the task-by-task
    results of the reference's CPU backend; mixed batches keep the reference's loop."""
    import lightworks_b200
    from lightworks.emulator import Backend

    p = lw.Parameter(0.1)
    circ = lw.PhotonicCircuit(6)
    for i in range(5):
        circ.bs(i, reflectivity=p)
        circ.ps(i, 0.3 * i)
    circ.add(lw.Unitary(lw.random_unitary(6, seed=3)))
    values = [0.1, 0.25, 0.4, 0.55, 0.7, 0.85]
    mk = lambda: lw.Batch(lw.Simulator, task_args=[[circ], [lw.State([1, 0, 1, 0, 1, 0])]],  # noqa: E731
                          parameters={p: values})
    ctx = lightworks_b200.default_context()
    b = Backend("permanent")
    l0 = ctx.launch_count()
    got = b.run(mk())
    launches = ctx.launch_count() - l0
    ref = Backend("permanent_cpu").run(mk())
    assert len(got) == len(ref) == len(values)
    assert launches <= 2, launches  # fock_basis enumeration + ONE permanent launch for the six unitaries
    for g, r in zip(got, ref):
        assert g.outputs == r.outputs and g.inputs == r.inputs
        assert _close(g.array, r.array)
    assert not _close(got[0].array, got[-1].array)  # the parameter really changed the unitary
    # explicit, differing outputs per task -> two groups; an Analyzer in the batch -> the reference's loop
    batch = lw.Batch()
    u = lw.Unitary(lw.random_unitary(4, seed=1))
    batch.add(lw.Simulator(u, lw.State([1, 0, 1, 0]), lw.State([0, 1, 0, 1])))
    batch.add(lw.Simulator(u, lw.State([1, 0, 1, 0]), lw.State([1, 1, 0, 0])))
    batch.add(lw.Simulator(u, lw.State([1, 0, 1, 0])))
    got = b.run(batch)
    ref = Backend("permanent_cpu").run(batch)
    for g, r in zip(got, ref):
        assert g.outputs == r.outputs and _close(g.array, r.array)
    mixed = lw.Batch()
    mixed.add(lw.Simulator(u, lw.State([1, 0, 1, 0])))
    mixed.add(lw.Analyzer(u, lw.State([1, 0, 1, 0])))
    got = b.run(mixed)
    assert [type(g).__name__ for g in got] == ["LazySimulationResult"] * 2
    assert got[1].result_type == "probability"

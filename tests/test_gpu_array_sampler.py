"""
Array-native Sampler on the GPU (lightworks_b200/array_sampler.py): the device selection mask of
lwb_basis_sample_selected against the host mask, the device regime against the exact distribution, and BASELINE
config 3 (24 modes, 10 photons, 92.6 M outputs) through the UNMODIFIED SDK: Backend("permanent").run(Sampler(...)).
"""
import time

import numpy as np
import pytest

from oracle import oracle as O  # noqa: N812

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lb():
    import lightworks_b200

    return lightworks_b200


@pytest.mark.parametrize("backend", [0, 1])
def test_device_selection_mask_equals_host_mask(lb, backend):
    """Probability mass kept by the device mask == mass kept by the numpy mask on the same states, and every drawn
    state passes the selection; threshold detection, heralds, min_detection and two post-selection rules together."""
    from lightworks_b200 import array_sampler as asmp

    ctx = lb.default_context(0)
    m, n = 9, 5
    U = O.random_unitary(m, 21)  # noqa: N806
    ins = [1, 1, 0, 1, 0, 1, 0, 1, 0]
    S = lb.fock_count(m, n)  # noqa: N806
    if backend == 0:
        probs, states = ctx.perm_basis_probs(U, ins), ctx.fock_basis(m, n)
    else:
        probs = ctx.slos_basis_probs(U, ins)
        states = np.array([lb.slos_unrank(m, n, r) for r in range(S)], dtype=np.uint8)
    rng = np.random.default_rng(1)
    cases = [
        dict(thr=False, mind=0, heralds={}, rules=[]),
        dict(thr=False, mind=3, heralds={2: 1, 7: 0}, rules=[]),
        dict(thr=True, mind=2, heralds={0: 1}, rules=[((1, 2), (0, 1)), ((3, 5, 6), (1, 2))]),
        dict(thr=False, mind=0, heralds={4: 2}, rules=[((0,), (0,))]),
    ]
    for c in cases:
        u = rng.random(4000)
        idx, total = ctx.basis_sample_selected(U, ins, u, 1e-9, backend, c["thr"], c["mind"], c["heralds"], c["rules"])
        st = np.minimum(states, 1) if c["thr"] else states
        kept = [j for j in range(m) if j not in c["heralds"]]
        # rules arrive in full-circuit modes already (the host maps them), so build the host mask the same way
        ok = np.ones(S, dtype=bool)
        for mode, cnt in c["heralds"].items():
            ok &= st[:, mode] == cnt
        ok &= st[:, kept].sum(axis=1) >= c["mind"]
        for modes, counts in c["rules"]:
            ok &= np.isin(st[:, list(modes)].sum(axis=1), counts)
        ok &= probs > 1e-9
        assert total == pytest.approx(probs[ok].sum(), rel=1e-12)
        assert ok[idx.astype(np.int64)].all()
        # frequencies follow the renormalised masked distribution (loose 6-sigma bound on the largest entry)
        top = int(np.argmax(np.where(ok, probs, 0)))
        p_top = probs[top] / probs[ok].sum()
        assert abs(np.mean(idx == top) - p_top) < 6 * np.sqrt(p_top * (1 - p_top) / len(u)) + 1e-3
    sel = asmp.Selection(m, {2: 1, 7: 0}, 3, type("P", (), {"rules": []})(), True)
    assert sel.kept == [0, 1, 3, 4, 5, 6, 8]


@pytest.mark.reference
def test_device_regime_matches_host_regime_statistically(lb, monkeypatch):
    """Force the device regime on a distribution small enough for the exact host regime and compare the sampled
    frequencies of a heralded, post-selected, threshold-detected Sampler (both sampling modes)."""
    from oracle import refshim

    lw = refshim.import_lightworks()
    from lightworks import PostSelection
    from lightworks.emulator import Backend, Detector

    from lightworks_b200 import array_sampler as asmp

    lb.install()
    try:
        circ = lw.PhotonicCircuit(10)
        circ.add(lw.Unitary(lw.random_unitary(10, seed=4)))
        circ.herald((0, 0), (1, 1))
        circ.herald((5, 4), (0, 0))
        ps = PostSelection()
        ps.add((0, 1), (0, 1))
        for mode, det in (("output", Detector(photon_counting=False)), ("input", Detector())):
            mk = lambda seed: lw.Sampler(circ, lw.State([1, 0, 1, 1, 0, 1, 0, 0]), 200_000, random_seed=seed,  # noqa: E731
                                         post_selection=ps, min_detection=2, detector=det, sampling_mode=mode)
            exact = Backend("permanent").run(mk(1))
            monkeypatch.setattr(asmp, "REGIME_A_MAX", 100)
            dev = Backend("permanent").run(mk(2))
            monkeypatch.setattr(asmp, "REGIME_A_MAX", 1 << 22)
            assert set(dev) <= set(exact) | set(dev)
            na, nb = sum(exact.values()), sum(dev.values())
            if mode == "output":
                assert na == nb == 200_000
            keys = set(exact) | set(dev)
            tv = 0.5 * sum(abs(exact.get(k, 0) / na - dev.get(k, 0) / nb) for k in keys)
            # two empirical distributions of N draws over K outcomes differ by about sqrt(K / (pi N)) in total variation
            assert tv < 1.5 * np.sqrt(len(keys) / (np.pi * min(na, nb))) + 0.005, (mode, tv, len(keys))
            assert abs(na - nb) < 6 * np.sqrt(200_000) + 1
    finally:
        lb.uninstall()


@pytest.mark.reference
def test_c3_sampler_through_the_unmodified_sdk(lb):
    """BASELINE config 3 as a user writes it: 24-mode Haar circuit, 10 photons, 10^6 samples, all 92 561 040 output
    states behind the draw.  The reference cannot run this (2 h, > 20 GB of State objects)."""
    from oracle import refshim

    lw = refshim.import_lightworks()
    from lightworks.emulator import Backend

    lb.install()
    try:
        circ = lw.Unitary(lw.random_unitary(24, seed=3))
        task = lw.Sampler(circ, lw.State([1] * 10 + [0] * 14), 1_000_000, random_seed=7)
        backend = Backend("permanent")
        backend.run(lw.Sampler(circ, lw.State([1] * 10 + [0] * 14), 1000, random_seed=1))  # warm-up (plans, buffers)
        t0 = time.perf_counter()
        res = Backend("permanent").run(task)
        dt = time.perf_counter() - t0
        assert sum(res.values()) == 1_000_000
        assert all(s.n_photons == 10 and len(s) == 24 for s in list(res)[:100])
        assert "lazy" in repr(task.probability_distribution)
        print(f"C3 Sampler through Backend('permanent').run: {dt:.3f} s, {len(res)} distinct outcomes")
        assert dt < 20
        # heralded + post-selected variant: the selection runs as a device mask over the whole basis
        circ2 = lw.PhotonicCircuit(24)
        circ2.add(lw.Unitary(lw.random_unitary(24, seed=3)))
        circ2.herald((0, 0), (1, 1))
        res2 = Backend("slos").run(lw.Sampler(circ2, lw.State([1] * 9 + [0] * 14), 100_000, random_seed=3, min_detection=9))
        assert sum(res2.values()) == 100_000 and all(len(s) == 23 and s.n_photons == 9 for s in list(res2)[:100])
    finally:
        lb.uninstall()

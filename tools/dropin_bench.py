"""End-to-end timings of unmodified Lightworks tasks on the B200 backends vs the reference CPU backends
(needs an importable reference: baseline/_ref or /root/reference).  python tools/dropin_bench.py"""
import json
import sys
import time

sys.path.insert(0, ".")
from oracle import refshim  # noqa: E402

lw = refshim.import_lightworks()
import lightworks_b200  # noqa: E402

lightworks_b200.install()
from lightworks.emulator import Backend  # noqa: E402


def timeit(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        best = min(best, time.perf_counter() - t0)
    return best, r


rows = []
for name, m, n, seed in [("C1", 6, 3, 1), ("C2", 12, 6, 2), ("m14n7", 14, 7, 7)]:
    circ = lw.Unitary(lw.random_unitary(m, seed=seed))
    st = lw.State([1] * n + [0] * (m - n))
    for task_name, mk, backends in [
        ("Simulator", lambda: lw.Simulator(circ, st), ("permanent",)),
        ("Analyzer", lambda: lw.Analyzer(circ, st), ("permanent",)),
        ("Sampler", lambda: lw.Sampler(circ, st, 10000, random_seed=1), ("permanent", "slos")),
    ]:
        for b in backends:
            if name == "m14n7" and task_name != "Sampler":
                reps_cpu = 1
            else:
                reps_cpu = 2
            # fresh Backend objects so the Sampler cache never short-circuits the distribution
            t_gpu, _ = timeit(lambda: Backend(b).run(mk()), 3)
            t_cpu, _ = timeit(lambda: Backend(b + "_cpu").run(mk()), reps_cpu)
            rows.append({"config": name, "modes": m, "photons": n, "task": task_name, "backend": b,
                         "b200_s": t_gpu, "reference_cpu_s": t_cpu, "speedup": t_cpu / t_gpu})
            print(rows[-1], flush=True)
# BASELINE config 3 through the unmodified SDK (array-native runner, distribution kept on the device)
circ = lw.Unitary(lw.random_unitary(24, seed=3))
st = lw.State([1] * 10 + [0] * 14)
Backend("permanent").run(lw.Sampler(circ, st, 1000, random_seed=1))
for b in ("permanent", "slos"):
    for n_smp in (10_000, 1_000_000):
        t, res = timeit(lambda: Backend(b).run(lw.Sampler(circ, st, n_smp, random_seed=2)), 2)
        rows.append({"config": "C3", "modes": 24, "photons": 10, "task": f"Sampler({n_smp} samples)", "backend": b,
                     "b200_s": t, "reference_cpu_s": None, "distinct_outcomes": len(res),
                     "note": "reference: ~2 h and > 20 GB of State objects (SURVEY.md 8d), not run"})
        print(rows[-1], flush=True)
print(json.dumps(rows))

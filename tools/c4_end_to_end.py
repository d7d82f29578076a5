"""BASELINE config 4: 16-mode Haar circuit (seed 4), 8 photons |1^8,0^8>,
Source(purity=0.99, brightness=0.9, indistinguishability=0.95) -> output distribution of the imperfect source.

Phases: native source statistics (lwb_source_statistics, host) -> pdist_calc on the device (255 unique sub-input
distributions through the permanent kernels, convolutions grouped by the indistinguishable photon group).
The reference needs 131 s for the first phase alone and hours for the second."""
import sys
import time

sys.path.insert(0, ".")
import lightworks_b200 as lb  # noqa: E402
from lightworks_b200.convolution import pdist_calc  # noqa: E402
from oracle import oracle as O  # noqa: E402, N812

m, n = 16, 8
U = O.random_unitary(m, 4)
circ = lb.CompiledCircuitView(U, m)
t0 = time.perf_counter()
annotated, entries = lb.source_statistics([1] * n + [0] * (m - n), purity=0.99, brightness=0.9, indistinguishability=0.95)
t1 = time.perf_counter()
print(f"source statistics: {len(entries)} annotated inputs in {t1 - t0:.2f} s (native, host)", flush=True)
b = lb.PermanentBackend()
small = dict(entries[:50])
pdist_calc(circ, small, b)  # warm-up: context, kernels
t2 = time.perf_counter()
res = pdist_calc(circ, dict(entries), b)
t3 = time.perf_counter()
vals = res.vals_array
print(f"pdist_calc on the device: {t3 - t2:.2f} s, {res.n_convolutions} convolutions, {len(vals)} output states, "
      f"sum = {vals.sum():.12f}, by photon number: "
      f"{ {int(k): float(vals[res.keys_array.sum(axis=1) == k].sum()) for k in range(0, 12)} }", flush=True)
t4 = time.perf_counter()
counts = lb.sample_source_outputs(circ, [1] * n + [0] * (m - n), 1_000_000, b, purity=0.99, brightness=0.9,
                                  indistinguishability=0.95, detector_efficiency=0.9, seed=1)
t5 = time.perf_counter()
tot = sum(counts.values())
mean_detected = sum(c * sum(k.s) for k, c in counts.items()) / tot
print(f"noisy Sampler on the device (source + convolution + 10^6 samples + detector efficiency 0.9): {t5 - t4:.2f} s, "
      f"{len(counts)} distinct outcomes, mean detected photons {mean_detected:.4f}", flush=True)

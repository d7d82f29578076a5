"""Source-statistics combination on the device (lwb_dist_*): C4-shaped workloads.

Annotated inputs have the structure Source._build_statistics produces for indistinguishability < 1 and
brightness < 1 (components/source.py:228-354): a subset G of the photons shares label 0 (they interfere), every
other PRESENT photon carries its own label, the rest are lost at the source -> 3^n inputs for n photons.
Weights are the corresponding binomial products.  The oracle (pure-Python restatement of the reference loops) is
timed on the smallest case only.
"""
import itertools
import sys
import time

sys.path.insert(0, ".")
import lightworks_b200 as lb  # noqa: E402
from lightworks_b200.convolution import pdist_calc  # noqa: E402
from oracle import oracle as O  # noqa: E402, N812


class Circ:
    def __init__(self, U, n_modes):  # noqa: N803
        self.U_full, self.n_modes, self.loss_modes = U, n_modes, U.shape[0] - n_modes


def inputs_for(m, n, indist=0.95, bright=0.9):
    out = {}
    for assign in itertools.product((0, 1, 2), repeat=n):  # 0 lost, 1 in the indistinguishable group, 2 own label
        modes = [[] for _ in range(m)]
        w = 1.0
        lab = 1
        for ph, a in enumerate(assign):
            if a == 0:
                w *= 1 - bright
            elif a == 1:
                w *= bright * indist
                modes[ph].append(0)
            else:
                w *= bright * (1 - indist)
                modes[ph].append(lab)
                lab += 1
        key = tuple(tuple(x) for x in modes)
        out[key] = out.get(key, 0.0) + w
    return out


b = lb.PermanentBackend()
for (m, n, seed, run_oracle) in [(8, 4, 1, True), (12, 6, 2, False), (16, 8, 4, False)]:
    U = O.random_unitary(m, seed)  # noqa: N806
    circ = Circ(U, m)
    inputs = inputs_for(m, n)
    pdist_calc(circ, dict(list(inputs.items())[:3]), b)  # warm-up (context, kernels)
    dt = 1e30
    for _ in range(2):  # best of two: the first pass also pays CUDA's lazy kernel loading
        t0 = time.perf_counter()
        res = pdist_calc(circ, inputs, b)
        dt = min(dt, time.perf_counter() - t0)
    print(f"m={m} n={n}: {len(inputs)} annotated inputs -> {len(res)} outputs, sum={res.vals_array.sum():.12f}, "
          f"{res.n_convolutions} convolutions, {dt * 1e3:.1f} ms on the device path")
    if run_oracle:
        t0 = time.perf_counter()
        ref = O.py_annotated_pdist_calc(circ, inputs, O.pdist_func("permanent"))
        dt_o = time.perf_counter() - t0
        got = {tuple(int(v) for v in k): float(p) for k, p in zip(res.keys_array, res.vals_array)}
        err = max(abs(got[k] - v) / v for k, v in ref.items())
        print(f"   oracle (pure-Python restatement of probability_distribution.py:78-164, 1 core): {dt_o * 1e3:.1f} ms, "
              f"max rel diff {err:.2e}, same states: {set(got) == set(ref)}")

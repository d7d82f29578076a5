"""
Generate tests/golden/ref_convolve.json by running the UNMODIFIED reference's pdist_calc /
annotated_state_pdist_calc (lightworks/emulator/simulation/probability_distribution.py:25-164) on the inputs its
own Source model produces, for both reference CPU backends.  Build container only:

    python -m oracle.make_golden_convolve

Every record holds: U_full (re, im lists), n_modes, loss_modes, backend, the `inputs` dict handed to pdist_calc
(State or AnnotatedState lists with their probabilities, in dict order) and the returned distribution
(keys / vals in dict order).  Test infrastructure only.
"""
from __future__ import annotations

import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "ref_convolve.json")


def main() -> int:
    from oracle import refshim

    lw = refshim.import_lightworks()
    from lightworks import PhotonicCircuit, State, Unitary, convert, random_unitary
    from lightworks.emulator import Source
    from lightworks.emulator.backends.permanent import PermanentBackend
    from lightworks.emulator.backends.slos import SLOSBackend
    from lightworks.emulator.simulation.probability_distribution import pdist_calc

    def lossy_circuit():  # tests/emulator/sampler_test.py:387-430
        c = PhotonicCircuit(4)
        c.bs(0, loss=convert.db_loss_to_decimal(1))
        c.bs(2, loss=convert.db_loss_to_decimal(2))
        c.ps(1, 0.3, loss=convert.db_loss_to_decimal(0.5))
        c.ps(3, 0.3, loss=convert.db_loss_to_decimal(0.5))
        c.bs(1, loss=convert.db_loss_to_decimal(1))
        c.bs(2, loss=convert.db_loss_to_decimal(2))
        c.ps(1, 0.3, loss=convert.db_loss_to_decimal(0.5))
        return c

    scenarios = [
        ("sampler_test.py:352-362 imperfect source", Unitary(random_unitary(4, seed=20)), [1, 0, 1, 0],
         Source(purity=0.9, brightness=0.9, indistinguishability=0.9)),
        ("sampler_test.py:375-385 2 photons imperfect", Unitary(random_unitary(4, seed=20)), [0, 2, 0, 0],
         Source(purity=0.9, brightness=0.9, indistinguishability=0.9)),
        ("sampler_test.py:409-430 lossy imperfect", lossy_circuit(), [1, 0, 1, 0],
         Source(purity=0.9, brightness=0.9, indistinguishability=0.9)),
        ("lossy, brightness only (State inputs, vacuum rule :70-73)", lossy_circuit(), [1, 0, 1, 0], Source(brightness=0.8)),
        ("haar m6 seed11, 3 photons, partial distinguishability", Unitary(random_unitary(6, seed=11)), [1, 1, 1, 0, 0, 0],
         Source(purity=0.95, brightness=0.9, indistinguishability=0.8)),
        ("haar m5 seed12, bunched input, distinguishability only", Unitary(random_unitary(5, seed=12)), [2, 0, 1, 0, 1],
         Source(indistinguishability=0.7)),
        ("perfect source (single State input)", Unitary(random_unitary(5, seed=13)), [1, 0, 1, 0, 1], Source()),
    ]
    records = []
    for name, circuit, in_state, source in scenarios:
        compiled = circuit._build()
        inputs = source._build_statistics(State(in_state))
        for bname, bk in (("permanent", PermanentBackend()), ("slos", SLOSBackend())):
            dist = pdist_calc(compiled, inputs, bk.full_probability_distribution)
            first = next(iter(inputs))
            annotated = type(first).__name__ == "AnnotatedState"
            records.append({
                "scenario": name, "backend": bname, "annotated": annotated,
                "n_modes": compiled.n_modes, "loss_modes": compiled.loss_modes,
                "U_re": compiled.U_full.real.tolist(), "U_im": compiled.U_full.imag.tolist(),
                "inputs": [[k.s, float(p)] for k, p in inputs.items()],
                "keys": [k.s for k in dist], "vals": [float(v) for v in dist.values()],
            })
            print(f"{name} [{bname}]: {len(inputs)} inputs ({'annotated' if annotated else 'state'}) -> {len(dist)} outputs, "
                  f"sum {sum(dist.values()):.12f}")
    with open(OUT, "w") as f:
        json.dump({"reference_version": getattr(lw, "__version__", "2.3.3"),
                   "generator": "python -m oracle.make_golden_convolve", "records": records}, f)
    print("bytes:", os.path.getsize(OUT))
    return 0


if __name__ == "__main__":
    sys.exit(main())

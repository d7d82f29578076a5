"""
Generate tests/golden/ref_source.json: Source._build_statistics (lightworks/emulator/components/source.py:134-161) of
the UNMODIFIED reference for a set of source settings and target states.  Build container only:

    python -m oracle.make_golden_source

Test infrastructure only.
"""
from __future__ import annotations

import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "ref_source.json")

CASES = [
    ([1, 0, 1, 0], dict(purity=0.9, brightness=0.9, indistinguishability=0.9)),
    ([0, 2, 0, 0], dict(purity=0.9, brightness=0.9, indistinguishability=0.9)),
    ([1, 0, 1, 0], dict(brightness=0.8)),
    ([1, 1, 1, 0, 0, 0], dict(purity=0.95, brightness=0.9, indistinguishability=0.8)),
    ([2, 0, 1, 0, 1], dict(indistinguishability=0.7)),
    ([1, 0, 1, 0, 1], dict()),
    ([0, 0, 0], dict(purity=0.9)),
    ([0, 0, 0], dict(brightness=0.5)),
    ([3, 1], dict(purity=0.8, brightness=0.7, indistinguishability=0.6, probability_threshold=1e-4)),
    ([1, 1, 0, 1, 1], dict(purity=0.99, brightness=0.9, indistinguishability=0.95)),
    ([1, 2, 0, 0, 1], dict(purity=1, brightness=0.6, indistinguishability=0.5)),
    ([1, 1, 1, 1, 1, 0, 0, 0], dict(purity=0.99, brightness=0.9, indistinguishability=0.95, probability_threshold=1e-7)),
    ([0, 1, 0, 0, 2, 0], dict(purity=0.7, brightness=1, indistinguishability=1)),
]


def main() -> int:
    from oracle import refshim

    lw = refshim.import_lightworks()
    from lightworks import State
    from lightworks.emulator import Source

    records = []
    for state, kw in CASES:
        src = Source(**kw)
        dist = src._build_statistics(State(state))
        first = next(iter(dist))
        records.append({"state": state, "source": kw, "annotated": type(first).__name__ == "AnnotatedState",
                        "keys": [k.s for k in dist], "vals": [float(v) for v in dist.values()]})
        print(state, kw, "->", len(dist), "inputs")
    with open(OUT, "w") as f:
        json.dump({"reference_version": getattr(lw, "__version__", "2.3.3"),
                   "generator": "python -m oracle.make_golden_source", "records": records}, f)
    print("bytes:", os.path.getsize(OUT))
    return 0


if __name__ == "__main__":
    sys.exit(main())

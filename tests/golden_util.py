"""Loader for tests/golden/ref_calls.{json,npz} (written by `python -m oracle.make_golden`)."""
import json
import os

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Golden:
    def __init__(self):
        with open(os.path.join(_DIR, "ref_calls.json")) as f:
            meta = json.load(f)
        self.records = meta["records"]
        self.literals = meta["literals"]
        self._npz = np.load(os.path.join(_DIR, "ref_calls.npz"))

    def arr(self, name):
        return self._npz[name]

    def of_kind(self, kind):
        return [r for r in self.records if r["kind"] == kind]


_golden = None


def golden() -> Golden:
    global _golden
    if _golden is None:
        _golden = Golden()
    return _golden


def rel_close(a, b, rel, abs_tol=0.0):
    a = np.asarray(a)
    b = np.asarray(b)
    return bool(np.all(np.abs(a - b) <= rel * np.abs(b) + abs_tol))

"""
Stand-in for `thewalrus` (pinned ==0.20.0 by the reference, not installed here).
`perm` follows thewalrus/_permanent.py as restated in oracle/lw_oracle.c (lwo_perm):
type/shape/NaN checks, closed forms for n<=3, Glynn/BBFG Gray-code loop otherwise.
Test infrastructure only.
"""
from __future__ import annotations

import numpy as np

from .. import oracle as _oracle


def perm(A, method="bbfg"):  # noqa: N803
    if not isinstance(A, np.ndarray):
        raise TypeError("Input matrix must be a NumPy array.")
    if A.ndim != 2 or A.shape[0] != A.shape[1]:
        raise ValueError("Input matrix must be square.")
    if np.isnan(A).any():
        raise ValueError("Input matrix must not contain NaNs.")
    if method != "bbfg":
        raise NotImplementedError("only the default bbfg method is restated")
    if A.shape[0] == 0:
        return A.dtype.type(1.0)
    return _oracle.perm(A)

"""
Minimal stand-ins for the two reference value types the hot path touches, used only when the
`lightworks` package is not importable (e.g. on a bare GPU box).  With lightworks present,
`lightworks_b200.install()` makes the backends return real `lightworks.State` objects instead.

  State               mirrors lightworks/sdk/state/state.py:27-140 (list wrapper, eq by content)
  CompiledCircuitView mirrors the fields of CompiledPhotonicCircuit that the backends read
                      (sdk/circuit/photonic_compiler.py:46-80): U_full, n_modes, loss_modes
"""
from __future__ import annotations

import numpy as np


class State:
    __slots__ = ("_s",)

    def __init__(self, state):
        self._s = state if isinstance(state, list) else list(state)

    @property
    def s(self) -> list[int]:
        return list(self._s)

    @property
    def n_photons(self) -> int:
        return sum(self._s)

    @property
    def n_modes(self) -> int:
        return len(self._s)

    def __str__(self) -> str:
        return "|" + ",".join(str(v) for v in self._s) + ">"

    def __repr__(self) -> str:
        return f"lightworks_b200.State({self})"

    def __add__(self, other: "State") -> "State":
        if not isinstance(other, State):
            raise TypeError("Addition only supported between states.")
        return State(self._s + other._s)

    def __eq__(self, other) -> bool:
        return isinstance(other, State) and self._s == other._s

    def __hash__(self) -> int:
        # equal states hash equally, which is all a dict needs; the reference hashes the ket string (state.py:95-96),
        # ten times slower for the 3 x 10^5 distinct outcomes of a noisy 10^6-sample run
        return hash(tuple(self._s))

    def __len__(self) -> int:
        return len(self._s)

    def __iter__(self):
        return iter(list(self._s))

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            return State(self._s[idx])
        return self._s[idx]


class CompiledCircuitView:
    """What the backends need from a compiled circuit: the full unitary (loss modes appended as
    extra rows/columns) and how many of its modes are real."""

    def __init__(self, U_full, n_modes: int | None = None):  # noqa: N803
        self.U_full = np.ascontiguousarray(U_full, dtype=np.complex128)
        if self.U_full.ndim != 2 or self.U_full.shape[0] != self.U_full.shape[1]:
            raise ValueError("U_full must be square")
        self.n_modes = self.U_full.shape[0] if n_modes is None else int(n_modes)
        self.loss_modes = self.U_full.shape[0] - self.n_modes
        if self.loss_modes < 0:
            raise ValueError("n_modes larger than the unitary")

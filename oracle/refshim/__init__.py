"""
oracle.refshim -- import shims that let the UNMODIFIED reference (Aegiq/lightworks
v2.3.3 at /root/reference) be imported in the build container.

TEST INFRASTRUCTURE ONLY (see oracle/lw_oracle.c header).  It exists to (1) pin the
oracle restatement against the real reference and (2) generate tests/golden/*.  It is
never imported by the product package and never runs on the GPU box
(/root/reference does not exist there).

What is shimmed and why:
  * `multimethod`  -- not installed; a small annotation dispatcher with the semantics
    the reference relies on (backends/backend.py:90-102, fock_backend.py:49-68,
    simulation/probability_distribution.py:25,78, ...).
  * `matplotlib`, `IPython`, `drawsvg` -- plotting only; stub modules.
  * `thewalrus.perm` -- pinned thewalrus==0.20.0 is not installed (no network); its
    algorithm is restated in oracle/lw_oracle.c (lwo_perm) and called through ctypes.
"""
from __future__ import annotations

import importlib
import importlib.machinery
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _find_reference() -> str:
    """/root/reference in the build container; on the GPU box the pip-installed copy under
    baseline/_ref (git-ignored, created by __graft_entry__.build(); it travels with gpurun)."""
    for cand in (os.environ.get("LW_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "lightworks")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_reference()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "lightworks"))


class _StubClass:
    """Stand-in for plotting classes used in annotations (e.g. `plt.Axes | NDArray`)."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return _StubClass()

    def __call__(self, *a, **k):
        return _StubClass()


class _StubModule(types.ModuleType):
    def __init__(self, name):
        super().__init__(name)
        self.__path__ = []  # behaves as a package
        self.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=True)

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        full = f"{self.__name__}.{name}"
        if name[0].isupper():
            obj = type(name, (_StubClass,), {})
        elif full in sys.modules:
            obj = sys.modules[full]
        else:
            obj = _StubModule(full)
            sys.modules[full] = obj
        setattr(self, name, obj)
        return obj

    def __call__(self, *a, **k):
        return _StubClass()


_installed = False


def install() -> None:
    """Inject the shims into sys.modules and put the reference on sys.path."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    from . import _multimethod, _thewalrus

    sys.modules.setdefault("multimethod", _multimethod)
    sys.modules.setdefault("thewalrus", _thewalrus)
    for name in (
        "matplotlib",
        "matplotlib.pyplot",
        "matplotlib.figure",
        "matplotlib.patches",
        "matplotlib.colors",
        "matplotlib.axes",
        "IPython",
        "IPython.display",
        "drawsvg",
    ):
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = _StubModule(name)
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)
    _installed = True


def import_lightworks():
    install()
    import lightworks  # noqa: PLC0415

    return lightworks

"""Minimal `multimethod` replacement (test infrastructure; see refshim/__init__.py)."""
from __future__ import annotations

import collections.abc as cabc
import functools
import inspect
import types
import typing


def _matches(arg, hint) -> bool:
    if hint is typing.Any or hint is inspect.Parameter.empty:
        return True
    origin = typing.get_origin(hint)
    if origin is typing.Union or origin is types.UnionType:
        return any(_matches(arg, h) for h in typing.get_args(hint))
    if origin is None:
        if isinstance(hint, type):
            return isinstance(arg, hint)
        return True
    if origin is cabc.Callable:
        return callable(arg)
    if not isinstance(origin, type):
        return True
    if not isinstance(arg, origin):
        return False
    args = typing.get_args(hint)
    if origin in (list, tuple, set, frozenset) and args and len(arg) > 0:
        first = next(iter(arg))
        return _matches(first, args[0])
    if origin is dict and args and len(arg) > 0:
        k, v = next(iter(arg.items()))
        return _matches(k, args[0]) and _matches(v, args[1])
    return True  # e.g. np.ndarray[Any, Any], Parameter[Any]


def _specificity(hints) -> int:
    return sum(0 if (h is typing.Any or h is inspect.Parameter.empty) else 1 for h in hints)


class multimethod:
    def __init__(self, func):
        self._impls = []
        self.__name__ = getattr(func, "__name__", "multimethod")
        self.__doc__ = getattr(func, "__doc__", None)
        self.register(func)

    def _hints(self, func):
        sig = inspect.signature(func)
        try:
            th = typing.get_type_hints(func)
        except Exception:
            th = {}
        names = [
            p.name
            for p in sig.parameters.values()
            if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)
        ]
        return names, [th.get(n, inspect.Parameter.empty) for n in names]

    def register(self, func):
        names, hints = self._hints(func)
        self._impls.append((func, names, hints))
        return func

    def _dispatch(self, args, kwargs):
        best = None
        for order, (func, names, hints) in enumerate(self._impls):
            if len(args) > len(names):
                continue
            ok = True
            for a, h in zip(args, hints):
                if not _matches(a, h):
                    ok = False
                    break
            if ok:
                for k, v in kwargs.items():
                    if k in names and not _matches(v, hints[names.index(k)]):
                        ok = False
                        break
            if ok:
                key = (_specificity(hints[: max(len(args), 1)] if not kwargs else hints), order)
                if best is None or key >= best[0]:
                    best = (key, func)
        if best is None:
            raise TypeError(f"{self.__name__}: no matching method for {[type(a) for a in args]}")
        return best[1]

    def __call__(self, *args, **kwargs):
        return self._dispatch(args, kwargs)(*args, **kwargs)

    def __get__(self, instance, owner=None):
        if instance is None:
            return self
        return functools.partial(self.__call__, instance)

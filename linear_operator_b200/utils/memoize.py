"""Per-object result cache (reference: utils/memoize.py:9-125): ``@cached(name=...)`` stores the return value in
``obj._memoize_cache`` keyed by (name, args, kwargs)."""
from __future__ import annotations

import functools
import pickle

from .errors import CachingError


def _key(name, args, kwargs):
    return (name, args, pickle.dumps(kwargs))


def add_to_cache(obj, name, val, *args, **kwargs):
    if not hasattr(obj, "_memoize_cache"):
        obj._memoize_cache = {}
    obj._memoize_cache[_key(name, args, kwargs)] = val
    return obj


def get_from_cache(obj, name, *args, **kwargs):
    try:
        return obj._memoize_cache[_key(name, args, kwargs)]
    except (AttributeError, KeyError):
        raise CachingError(f"Object does not have item {name} stored in cache.") from None


def pop_from_cache(obj, name, *args, **kwargs):
    try:
        return obj._memoize_cache.pop(_key(name, args, kwargs))
    except (AttributeError, KeyError):
        raise CachingError(f"Object does not have item {name} stored in cache.") from None


def _is_in_cache(obj, name, *args, **kwargs):
    return hasattr(obj, "_memoize_cache") and _key(name, args, kwargs) in obj._memoize_cache


def _is_in_cache_ignore_args(obj, name):
    return hasattr(obj, "_memoize_cache") and any(k[0] == name for k in obj._memoize_cache)


_is_in_cache_ignore_all_args = _is_in_cache_ignore_args


def clear_cache_hook(module, *args, **kwargs):
    module._memoize_cache = {}


def cached(method=None, name=None, ignore_args=False):
    if method is None:
        return functools.partial(cached, name=name, ignore_args=ignore_args)
    cache_name = name if name is not None else method

    @functools.wraps(method)
    def wrapper(self, *args, **kwargs):
        a, kw = ((), {}) if ignore_args else (args, kwargs)
        if not _is_in_cache(self, cache_name, *a, **kw):
            add_to_cache(self, cache_name, method(self, *args, **kwargs), *a, **kw)
        return get_from_cache(self, cache_name, *a, **kw)

    return wrapper

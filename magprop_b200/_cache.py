"""Small LRU of CUDA likelihood handles keyed by (model spec, grid, data bytes, prior)."""
from __future__ import annotations

import collections
import ctypes as C
import hashlib

import numpy as np

from . import _capi as A
from .engine import Likelihood, time_grid

_MAX = 16
_cache: "collections.OrderedDict[tuple, Likelihood]" = collections.OrderedDict()


def _digest(*arrays) -> str:
    h = hashlib.blake2b(digest_size=16)
    for a in arrays:
        if a is None:
            h.update(b"-")
        else:
            a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
            h.update(str(a.shape).encode())
            h.update(a.tobytes())
    return h.hexdigest()


def get(spec: A.ModelSpec, GRBtype, x=None, y=None, yerr=None, lower=None, upper=None, device=0) -> Likelihood:
    grid = time_grid(GRBtype)  # raises ValueError for a bad GRBtype (magnetar/funcs.py:138-141)
    key = (bytes(spec), "S" if GRBtype == "S" else "L", _digest(x, y, yerr, lower, upper), device)
    lk = _cache.get(key)
    if lk is None:
        lk = Likelihood(spec, grid, x, y, yerr, lower, upper, device=device)
        _cache[key] = lk
        while len(_cache) > _MAX:
            _, old = _cache.popitem(last=False)
            old.close()
    else:
        _cache.move_to_end(key)
    return lk


def clear():
    while _cache:
        _, old = _cache.popitem()
        old.close()

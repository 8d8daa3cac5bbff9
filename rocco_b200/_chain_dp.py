"""Drop-in for the reference's ``rocco._chain_dp`` extension module (``_chain_dp.c:9-213``).

Same single entry, same positional signature, same exceptions; the work happens in
``rocco_solve_penalized_chain_f64`` (CUDA, ``csrc/chain.cu``).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib


def solve_penalized_chain(scores, switch_costs, selection_penalty, /):
    s = np.ascontiguousarray(scores, dtype=np.float64)
    c = np.ascontiguousarray(switch_costs, dtype=np.float64)
    if s.ndim != 1:
        raise ValueError("`scores` must be one-dimensional")
    if c.ndim != 1:
        raise ValueError("`switch_costs` must be one-dimensional")
    n = s.shape[0]
    if n <= 0:
        raise ValueError("`scores` cannot be empty")
    if n > 1 and c.shape[0] != n - 1:
        raise ValueError("`switch_costs` must have length len(scores) - 1")
    lib = _lib.load()
    _lib.require_device()
    solution = np.zeros(n, dtype=np.uint8)
    value = ctypes.c_double(0.0)
    count = ctypes.c_longlong(0)
    st = lib.rocco_solve_penalized_chain_f64(
        _lib.np_ptr(s), _lib.np_ptr(c) if n > 1 else None, n, float(selection_penalty),
        _lib.np_ptr(solution), ctypes.byref(value), ctypes.byref(count))
    _lib.check(st, "solve_penalized_chain")
    return solution, float(value.value), int(count.value)

"""Drop-in for the reference's ``rocco._baseline`` extension module (``_baseline.c:16-104``)."""
from __future__ import annotations

import numpy as np

from . import _lib


def crossfit_whittaker_baseline(values, penalty_lambda):
    v = np.ascontiguousarray(values, dtype=np.float64)
    if v.ndim not in (1, 2):
        raise ValueError("`values` must be one-dimensional or two-dimensional")
    out = np.zeros(v.shape, dtype=np.float64)
    if v.size == 0:
        return out
    lib = _lib.load()
    _lib.require_device()
    rows, cols = (1, v.shape[0]) if v.ndim == 1 else v.shape
    st = lib.rocco_crossfit_whittaker_baseline_matrix_f64(_lib.np_ptr(v), rows, cols, float(penalty_lambda), _lib.np_ptr(out))
    if st == _lib.ST_NOMEM:
        raise MemoryError()
    _lib.check(st, "crossfit_whittaker_baseline")
    return out

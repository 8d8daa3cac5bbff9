"""Drop-in for the reference's ``rocco._wls`` extension module (``_wls.c:8-162``): same keyword
signature, same 8-tuple, same exceptions."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib


def score_centered_wls(centered_matrix, lower_bound_z=1.0, prior_df=5.0, min_effect=None, spatial_window=31,
                       precision_floor_ratio=0.01):
    x = np.ascontiguousarray(centered_matrix, dtype=np.float64)
    if x.ndim != 2:
        raise ValueError("`centered_matrix` must be two-dimensional")
    use_min_effect = 0
    min_effect_ = 0.0
    if min_effect is not None:
        min_effect_ = float(min_effect)
        if min_effect_ < 0.0:
            min_effect_ = 0.0
        use_min_effect = 1
    m, n = x.shape
    outs = [np.zeros(n, dtype=np.float64) for _ in range(6)]     # mean raw prior moderated se scores
    lib = _lib.load()
    if m == 0 or n == 0:
        raise ValueError("Invalid centered-WLS inputs")
    _lib.require_device()
    total_df = ctypes.c_double(0.0)
    window = ctypes.c_int(0)
    st = lib.rocco_score_centered_wls_f64(
        _lib.np_ptr(x), m, n, float(lower_bound_z), float(prior_df), min_effect_, use_min_effect,
        int(spatial_window), float(precision_floor_ratio),
        *[_lib.np_ptr(o) for o in outs], ctypes.byref(total_df), ctypes.byref(window))
    if st == _lib.ST_NOMEM:
        raise MemoryError()
    if st == _lib.ST_INVALID:
        raise ValueError("Invalid centered-WLS inputs")
    if st == _lib.ST_NONFINITE:
        raise ValueError("EB scoring produced non-finite values")
    _lib.check(st, "score_centered_wls")
    mean, raw, prior, mod, se, scores = outs
    return scores, mean, raw, prior, mod, se, float(total_df.value), int(window.value)

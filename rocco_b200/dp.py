"""Host-side mirror of the reference's ``rocco/dp.py`` (same names, arguments, return types).

The chain DP and the budget search run on the GPU (``csrc/chain.cu``); this module only keeps the
reference's Python-level contract: ``dp.py:16-34`` objective_value, ``37-46`` build_switch_costs,
``49-86`` solve_penalized_chain, ``89-164`` calibrate_selection_penalty, ``167-228`` solve_chrom_exact.
"""
from __future__ import annotations

import ctypes
import logging
from typing import Dict, Optional, Tuple

import numpy as np

from . import _chain_dp, _lib

logger = logging.getLogger(__name__)


def _vector(values, what: str) -> np.ndarray:
    arr = np.ascontiguousarray(values, dtype=np.float64)
    if arr.ndim != 1:
        raise ValueError(f"`{what}` must be a one-dimensional array")
    return arr


def objective_value(solution: np.ndarray, scores: np.ndarray, switch_costs) -> float:
    r"""-s.z + c.|dz|: the unpenalised objective the reference reports (dp.py:16-34); a scalar cost applies to every
    adjacent pair."""
    z = np.asarray(solution, dtype=np.float64)
    gain = float(np.asarray(scores, dtype=np.float64) @ z)
    if z.shape[0] < 2:
        return -gain
    jumps = np.abs(z[1:] - z[:-1])
    cost = float(switch_costs) * float(jumps.sum()) if np.isscalar(switch_costs) \
        else float(np.asarray(switch_costs, dtype=np.float64) @ jumps)
    return cost - gain


def build_switch_costs(scores: np.ndarray, gamma: float = 1.0) -> np.ndarray:
    r"""The constant fragmentation cost between neighbouring bins, length n - 1 (dp.py:37-46).  The device entries take
    gamma as a scalar; this vector exists for callers of the reference API."""
    n = _vector(scores, "scores").shape[0]
    return np.full(max(n - 1, 0), float(gamma), dtype=np.float64)


def solve_penalized_chain(scores, switch_costs, selection_penalty: float) -> Tuple[np.ndarray, float, int]:
    r"""max_z sum (s_j - lambda) z_j - sum c_j |z_{j+1} - z_j|, ties -> fewer selected (dp.py:49-86).

    Returns (uint8 mask, penalized objective, selected count) with the reference's Python types."""
    mask, value, count = _chain_dp.solve_penalized_chain(
        np.ascontiguousarray(scores, dtype=np.float64), np.ascontiguousarray(switch_costs, dtype=np.float64),
        float(selection_penalty))
    return np.asarray(mask, dtype=np.uint8), float(value), int(count)


def calibrate_selection_penalty(scores, switch_costs, target_count: int, max_iter: int = 60
                                ) -> Tuple[float, np.ndarray, float, int]:
    r"""Bisection on DP solutions (dp.py:89-164), evaluated as a batched bisection tree on the GPU:
    the same (lower+upper)/2 lattice, several levels per launch."""
    scores_ = np.ascontiguousarray(scores, dtype=np.float64)
    switch_costs_ = np.ascontiguousarray(switch_costs, dtype=np.float64)
    n = scores_.shape[0]
    if scores_.ndim != 1 or switch_costs_.ndim != 1:
        raise ValueError("`scores` and `switch_costs` must be one-dimensional")
    if n == 0:
        raise ValueError("`scores` cannot be empty")
    if n > 1 and switch_costs_.shape[0] != n - 1:
        raise ValueError("`switch_costs` must have length len(scores) - 1")
    lib = _lib.load()
    _lib.require_device()
    solution = np.zeros(n, dtype=np.uint8)
    lam, val, cnt = ctypes.c_double(0.0), ctypes.c_double(0.0), ctypes.c_longlong(0)
    st = lib.rocco_calibrate_selection_penalty_f64(
        _lib.np_ptr(scores_), _lib.np_ptr(switch_costs_) if n > 1 else None, n, int(target_count), int(max_iter),
        ctypes.byref(lam), _lib.np_ptr(solution), ctypes.byref(val), ctypes.byref(cnt))
    _lib.check(st, "calibrate_selection_penalty")
    return float(lam.value), solution, float(val.value), int(cnt.value)


def solve_chrom_exact(
    scores: np.ndarray,
    budget: Optional[float] = None,
    gamma: float = 1.0,
    selection_penalty: Optional[float] = None,
    return_details: bool = False,
) -> Tuple[np.ndarray, float] | Tuple[np.ndarray, float, Dict[str, float]]:
    r"""Solve one chromosome with the exact penalized-chain DP (dp.py:167-228).

    ``scores`` may be a NumPy array or a CUDA ``torch.Tensor`` (float64); the latter skips the host
    round trip but NumPy arrays are still returned."""
    from .pipeline import solve_chromosomes

    res = solve_chromosomes([scores], budgets=[budget], gammas=[gamma], selection_penalties=[selection_penalty])[0]
    solution = res["solution"]
    if not return_details:
        return solution, res["objective"]
    return solution, res["objective"], {
        "penalized_objective": float(res["penalized_objective"]),
        "selected_count": int(res["selected_count"]),
        "selected_fraction": float(res["selected_count"] / solution.shape[0]),
        "selection_penalty": float(res["selection_penalty"]),
    }

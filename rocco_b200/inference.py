"""Host-side mirror of the scoring half of the reference's ``rocco/inference.py`` (lines 32-379).

Same function names, arguments, return types and exceptions.  ``score_loci_wls`` runs the whole
chain (log2p1 -> pilot offset -> cross-fit Whittaker baseline -> centered WLS) in one C-ABI call so the
reference's four m x n float64 temporaries (inference.py:325-341) never exist on the host; the stage
functions are kept for callers (and tests) that use them on their own.
"""
from __future__ import annotations

import ctypes
import logging
from typing import Any, Dict, Tuple

import numpy as np

from . import _baseline as _baseline_native
from . import _lib
from . import _wls as _wls_native

logger = logging.getLogger(__name__)


def _log_scale_wls_matrix(chrom_matrix: np.ndarray, pseudocount: float = 1.0) -> np.ndarray:
    matrix = np.asarray(chrom_matrix, dtype=np.float64)
    if np.any(~np.isfinite(matrix)):
        raise ValueError("`chrom_matrix` contains non-finite values")
    return np.log2(np.clip(matrix, 0.0, None) + float(pseudocount))


def _resolve_local_baseline_window(n_loci: int, target_window: int = 101) -> int:
    n_loci = int(n_loci)
    if n_loci < 25:
        return 0
    window = int(max(3, target_window))
    if window > n_loci:
        window = n_loci
    if (window % 2) == 0:
        window = window - 1 if window == n_loci else window + 1
    return int(max(0, window))


def _consenrich_whittaker_lambda(block_size: int) -> float:
    block = int(max(3, block_size))
    if (block % 2) == 0:
        block += 1
    w_hat = float(block) * 0.15915494
    return float(7.0 * (w_hat**4))


def _consenrich_crossfit_whittaker_baseline(y_vals: np.ndarray, block_size: int = 101) -> np.ndarray:
    y_arr = np.asarray(y_vals, dtype=np.float64)
    if y_arr.ndim != 1:
        raise ValueError("`y_vals` must be one-dimensional")
    window = _resolve_local_baseline_window(int(y_arr.size), target_window=block_size)
    if window == 0:
        return np.zeros_like(y_arr, dtype=np.float64)
    penalty_lambda = _consenrich_whittaker_lambda(window)
    return np.asarray(_baseline_native.crossfit_whittaker_baseline(y_arr, penalty_lambda=penalty_lambda), dtype=np.float64)


def _estimate_local_background_matrix(centered_matrix: np.ndarray, target_window: int = 101) -> tuple[np.ndarray, int, float]:
    matrix = np.asarray(centered_matrix, dtype=np.float64)
    if matrix.ndim != 2:
        raise ValueError("`centered_matrix` must be two-dimensional")
    window = _resolve_local_baseline_window(matrix.shape[1], target_window=target_window)
    if window == 0:
        return np.zeros_like(matrix, dtype=np.float64), 0, 0.0
    penalty_lambda = _consenrich_whittaker_lambda(window)
    local_baselines = np.asarray(
        _baseline_native.crossfit_whittaker_baseline(matrix, penalty_lambda=penalty_lambda), dtype=np.float64)
    if not np.all(np.isfinite(local_baselines)):
        raise ValueError("Local baseline fit produced non-finite values")
    return local_baselines, window, penalty_lambda


def _score_centered_wls_matrix(centered_matrix: np.ndarray, lower_bound_z: float = 1.0, prior_df: float = 5.0,
                               min_effect: float | None = None, spatial_window: int | None = None,
                               precision_floor_ratio: float = 0.01) -> tuple[np.ndarray, Dict[str, np.ndarray | float]]:
    centered = np.asarray(centered_matrix, dtype=np.float64)
    if centered.ndim != 2:
        raise ValueError("`centered_matrix` must be two-dimensional")
    if centered.shape[0] == 0 or centered.shape[1] == 0:
        raise ValueError("`centered_matrix` must be non-empty")
    precision_floor_ratio_ = float(max(precision_floor_ratio, 0.0))
    (scores_arr, mean_arr, raw_var_arr, prior_var_arr, moderated_var_arr, se_arr, total_df, resolved_window,
     ) = _wls_native.score_centered_wls(
        centered, lower_bound_z=float(lower_bound_z), prior_df=float(prior_df), min_effect=min_effect,
        spatial_window=31 if spatial_window is None else int(spatial_window),
        precision_floor_ratio=precision_floor_ratio_)
    se = np.asarray(se_arr, dtype=np.float64)
    mean = np.asarray(mean_arr, dtype=np.float64)
    scores = np.asarray(scores_arr, dtype=np.float64)
    details = {
        "mean": mean,
        "raw_variance": np.asarray(raw_var_arr, dtype=np.float64),
        "prior_variance": np.asarray(prior_var_arr, dtype=np.float64),
        "moderated_variance": np.asarray(moderated_var_arr, dtype=np.float64),
        "standard_error": se,
        "z_scores": mean / np.maximum(se, 1.0e-8),
        "min_effect": float(0.0 if min_effect is None else max(min_effect, 0.0)),
        "precision_floor_ratio": float(precision_floor_ratio_),
        "degrees_of_freedom": np.full(centered.shape[1], float(total_df), dtype=np.float64),
        "prior_spatial_window": float(resolved_window),
    }
    if (not np.all(np.isfinite(scores)) or not np.all(np.isfinite(details["mean"]))
            or not np.all(np.isfinite(details["raw_variance"])) or not np.all(np.isfinite(details["prior_variance"]))
            or not np.all(np.isfinite(details["moderated_variance"])) or not np.all(np.isfinite(details["standard_error"]))
            or not np.all(np.isfinite(details["z_scores"]))):
        raise ValueError("EB scoring produced non-finite values")
    return scores, details


def score_loci_wls(chrom_matrix: np.ndarray, lower_bound_z: float = 1.0, prior_df: float = 5.0,
                   min_effect: float | None = None, precision_floor_ratio: float = 0.01, low_memory: bool = False,
                   return_details: bool = False) -> np.ndarray | Tuple[np.ndarray, Dict[str, Any]]:
    r"""Score loci with an EB-moderated summary on baseline-corrected log signal (inference.py:302-379).

    float32 input is uploaded as float32 (half the host->device bytes) and widened on the device,
    which is exactly ``np.asarray(chrom_matrix, dtype=np.float64)`` of the reference."""
    matrix_in = np.asarray(chrom_matrix)
    if matrix_in.dtype != np.float32:
        matrix_in = np.asarray(matrix_in, dtype=np.float64)
    if matrix_in.ndim != 2:
        if np.any(~np.isfinite(matrix_in)):
            raise ValueError("`chrom_matrix` contains non-finite values")
        raise ValueError("`chrom_matrix` must be two-dimensional")
    if matrix_in.shape[0] == 0 or matrix_in.shape[1] == 0:
        raise ValueError("`chrom_matrix` must be non-empty")
    matrix_in = np.ascontiguousarray(matrix_in)
    m, n = matrix_in.shape
    lib = _lib.load()
    _lib.require_device()

    prm = _lib.ScoreParams()
    lib.rocco_b200_default_score_params(ctypes.byref(prm))
    prm.lower_bound_z = float(lower_bound_z)
    prm.prior_df = float(prior_df)
    prm.use_min_effect = 0 if min_effect is None else 1
    prm.min_effect = 0.0 if min_effect is None else max(float(min_effect), 0.0)
    prm.precision_floor_ratio = float(max(precision_floor_ratio, 0.0))

    names = ("scores", "mean", "raw_variance", "prior_variance", "moderated_variance", "standard_error")
    bufs = {k: np.zeros(n, dtype=np.float64) for k in (names if return_details else names[:1])}
    out = _lib.ScoreOutputs()
    for k, v in bufs.items():
        setattr(out, k, v.ctypes.data)
    centered = None
    if return_details:
        centered = np.zeros((m, n), dtype=np.float64)
        out.centered_matrix = centered.ctypes.data
    fn = lib.rocco_score_loci_wls_f32 if matrix_in.dtype == np.float32 else lib.rocco_score_loci_wls_f64
    st = fn(_lib.np_ptr(matrix_in), m, n, ctypes.byref(prm), ctypes.byref(out))
    if st == _lib.ST_NONFINITE:
        if np.any(~np.isfinite(matrix_in)):
            raise ValueError("`chrom_matrix` contains non-finite values")
        raise ValueError("Locus scoring produced non-finite values")
    _lib.check(st, "score_loci_wls")
    scores = bufs["scores"]
    if not return_details:
        return scores.astype(np.float64)
    se = bufs["standard_error"]
    details = {
        "input_scale": "log2p1",
        "local_baseline_window": int(out.baseline_window),
        "local_baseline_lambda": float(out.baseline_lambda),
        "mean": bufs["mean"],
        "raw_variance": bufs["raw_variance"],
        "prior_variance": bufs["prior_variance"],
        "moderated_variance": bufs["moderated_variance"],
        "standard_error": se,
        "z_scores": bufs["mean"] / np.maximum(se, 1.0e-8),
        "min_effect": float(0.0 if min_effect is None else max(min_effect, 0.0)),
        "precision_floor_ratio": float(max(precision_floor_ratio, 0.0)),
        "prior_spatial_window": int(out.resolved_spatial_window),
        "degrees_of_freedom": np.full(n, float(out.total_df), dtype=np.float64),
        "centered_matrix": centered.astype(np.float32 if low_memory else np.float64, copy=False),
    }
    return scores.astype(np.float64), details

"""ORACLE -- test infrastructure only (see oracle/oracle.py for the rules).

CPU restatement of the step either side of the hot path that SURVEY.md 8(f) ranks next:
the dependent-wild-bootstrap budget null and the automatic gamma
(/root/reference/rocco/inference.py:32-37, 446-681, 684-1148 and rocco.py:751-789).

Pinned by ``tests/test_oracle_pin.py::test_budget_*`` against golden vectors produced by the
real reference (``tests/golden/make_golden.py``): with the default stream factory below the
innovations are NumPy's PCG64 streams seeded exactly as the reference seeds them
(``base_seed + 104729 * (draw + 1)``, one generator per draw, samples in order), so every
number of the reference's details dict is reproduced, not just its distribution.
"""
from __future__ import annotations

import math

import numpy as np

from . import oracle as _o


# ------------------------------------------------------------------ small helpers
def robust_scale(values, floor: float = 1.0e-6) -> float:
    """1.4826 * MAD, floored (inference.py:32-37)."""
    v = np.asarray(values, dtype=np.float64)
    if v.size == 0:
        return float(floor)
    return float(max(1.4826 * np.median(np.abs(v - np.median(v))), floor))


def ess_max_lag(n_loci: int, dependence_lag_hint=None) -> int:
    """inference.py:504-517"""
    n = max(1, int(n_loci))
    base = min(n, 101) if dependence_lag_hint is None else max(1, min(n, int(dependence_lag_hint)))
    return int(min(n - 1, max(16, 4 * base)))


def bootstrap_bandwidth(n_loci: int, dependence_lag_hint=None) -> int:
    """inference.py:520-530"""
    n = max(1, int(n_loci))
    if n <= 1:
        return 1
    want = round(n ** (1.0 / 3.0)) if dependence_lag_hint is None else int(dependence_lag_hint)
    return int(min(n - 1, max(8, want)))


def bartlett_kernel(bandwidth: int) -> np.ndarray:
    """Triangular taps on [-b, b], unit L2 norm (inference.py:533-541)."""
    b = max(1, int(bandwidth))
    taps = np.maximum(1.0 - np.abs(np.arange(-b, b + 1, dtype=np.float64)) / float(b + 1), 0.0)
    return taps / np.sqrt(np.sum(taps * taps))


def effective_sample_size(values, max_lag: int):
    """n / tau_int with Geyer's initial-positive-sequence cut (inference.py:446-501).
    The autocovariances are direct lag products here (the reference takes them from an FFT; same numbers to rounding)."""
    v = np.asarray(values, dtype=np.float64)
    if v.ndim != 1:
        raise ValueError("`values` must be one-dimensional")
    n = int(v.size)
    if n < 4:
        return float(max(1, n)), 1.0, 0
    c = v - float(np.mean(v))
    var0 = float(np.mean(c * c))
    if not np.isfinite(var0) or var0 <= 1.0e-12:
        return float(n), 1.0, 0
    L = int(min(max(2, max_lag), n - 1))
    acov = np.array([np.dot(c[: n - k], c[k:]) / float(n - k) for k in range(L + 1)])
    if not np.isfinite(acov[0]) or acov[0] <= 1.0e-12:
        return float(n), 1.0, 0
    return geyer_tau(acov, n, L)


def geyer_tau(acov, n: int, L: int):
    rho = np.clip(np.asarray(acov[1:], dtype=np.float64) / float(acov[0]), -1.0, 1.0)
    tau, used = 1.0, 0
    for k in range(0, rho.size, 2):
        pair = float(rho[k]) + (float(rho[k + 1]) if k + 1 < rho.size else 0.0)
        if not np.isfinite(pair) or pair <= 0.0:
            break
        tau += 2.0 * pair
        used = int(min(L, k + 2))
    return float(np.clip(n / max(tau, 1.0), 1.0, n)), float(tau), used


# ------------------------------------------------------------------ multiplier field
def numpy_innovations(random_seed: int):
    """Stream factory reproducing the reference: draw d owns default_rng(seed + 104729 (d + 1)); samples are served in
    order from that one generator (inference.py:653, 656-662)."""
    state = {}

    def take(draw: int, sample: int, size: int) -> np.ndarray:
        if sample == 0:
            state[draw] = np.random.default_rng(int(random_seed) + 104729 * (int(draw) + 1))
        return state[draw].standard_normal(size)

    return take


def dependent_wild_weights(innovations, kernel) -> np.ndarray:
    """Valid-mode FIR of iid innovations with the Bartlett taps, then centred and scaled to unit (population) s.d.
    (inference.py:544-570; direct correlation instead of scipy.signal.fftconvolve -- same numbers to rounding).  The
    degenerate-scale fallback of the reference (Rademacher signs) cannot trigger for Gaussian innovations and n >= 2."""
    w = np.correlate(np.asarray(innovations, dtype=np.float64), np.asarray(kernel, dtype=np.float64), mode="valid")
    w = w - float(np.mean(w))
    s = float(np.std(w))
    if not np.isfinite(s) or s <= 1.0e-8:
        raise ValueError("degenerate multiplier field")
    return w / s


# ------------------------------------------------------------------ the null fit and the estimator
def fit_null_template(centered, kind="port", **score_kw):
    """Residual template  e~_ij = y_ij - max(mu_hat_j, 0)  (inference.py:684-716)."""
    y = np.asarray(centered, dtype=np.float64)
    scores, det = _o.score_centered_wls_matrix(y, kind=kind, **score_kw)
    pos = np.clip(np.asarray(det["mean"], dtype=np.float64), 0.0, None)
    return y - pos[None, :], scores.astype(np.float64), pos


def null_center_scale(null_scores):
    """Median of the fitted-null score field and the robust scale of its mirrored non-positive side
    (inference.py:768-783).  Returns (center, scale, negative_support_size)."""
    s = np.asarray(null_scores, dtype=np.float64)
    center = float(np.median(s))
    r = s - center
    neg = r[r <= 0.0]
    mag = np.abs(r) if neg.size == 0 else -neg
    if mag.size == 0:
        mag = np.zeros(1)
    return center, robust_scale(np.concatenate((-mag, mag))), int(mag.size)


def draw_statistics(scores, center: float, soft_scale: float, threshold: float):
    """The four per-draw means (inference.py:672-681)."""
    s = np.asarray(scores, dtype=np.float64)
    pos = np.clip(s - center, 0.0, None)
    return (float(np.mean(pos)), float(np.mean(pos / soft_scale)), float(np.mean(pos > 0.0)), float(np.mean(s > threshold)))


class Welford:
    """Running mean / M2 (inference.py:573-587)."""

    def __init__(self):
        self.n, self.mean, self.m2 = 0, 0.0, 0.0

    def add(self, x: float):
        self.n += 1
        d = float(x) - self.mean
        self.mean += d / float(self.n)
        self.m2 += d * (float(x) - self.mean)

    def sd(self) -> float:
        return float(math.sqrt(max(self.m2 / float(max(self.n - 1, 1)), 0.0)))

    def stderr(self) -> float:
        return float(math.sqrt(max(self.m2 / float(max(self.n - 1, 1)), 0.0) / float(max(self.n, 1))))


def stable_enough(w: Welford, min_draws: int, abs_tol: float, rel_tol: float) -> bool:
    """inference.py:590-603"""
    if w.n < max(2, int(min_draws)):
        return False
    return bool(w.stderr() <= max(abs_tol, rel_tol * max(abs(w.mean), 1.0e-6)))


def wild_bootstrap_score_null(centered, lower_bound_z=1.0, prior_df=5.0, min_effect=None, precision_floor_ratio=0.01,
                              observed_scores=None, dependence_lag_hint=None, num_null_draws=25, random_seed=0,
                              min_null_draws=None, stability_abs_tol=5.0e-3, stability_rel_tol=5.0e-2,
                              innovations=None, kind="port"):
    """inference.py:719-985 (single-process branch; the pool branch computes the same draws in batches)."""
    kw = dict(lower_bound_z=lower_bound_z, prior_df=prior_df, min_effect=min_effect, precision_floor_ratio=precision_floor_ratio)
    y = np.asarray(centered, dtype=np.float64)
    template, fitted, pos = fit_null_template(y, kind=kind, **kw)
    if observed_scores is None:
        observed = fitted
    else:
        observed = np.asarray(observed_scores, dtype=np.float64)
        if observed.shape[0] != y.shape[1]:
            raise ValueError("`observed_scores` must have the same number of loci as `centered_matrix`")
    ref_scores, _ = _o.score_centered_wls_matrix(template, kind=kind, **kw)
    center, scale, support = null_center_scale(ref_scores)
    if not np.isfinite(center) or not np.isfinite(scale):
        raise ValueError("Budget null fit produced non-finite values")
    soft = float(max(scale, 1.0e-6))
    threshold = float(center + 2.0 * scale)
    m, n = y.shape
    bw = bootstrap_bandwidth(n, dependence_lag_hint)
    taps = bartlett_kernel(bw)
    draws = int(max(1, num_null_draws))
    min_draws = int(min(draws, max(4, 8 if min_null_draws is None else min_null_draws)))
    take = numpy_innovations(random_seed) if innovations is None else innovations
    acc = [Welford() for _ in range(4)]                       # mass, units, fraction, tail occupancy
    for d in range(draws):
        boot = np.empty_like(template)
        for i in range(m):
            if n == 1:
                boot[i] = template[i] * 1.0
            else:
                boot[i] = template[i] * dependent_wild_weights(take(d, i, n + taps.size - 1), taps)
        sc, _ = _o.score_centered_wls_matrix(boot, kind=kind, **kw)
        for a, v in zip(acc, draw_statistics(sc, center, soft, threshold)):
            a.add(v)
        if stable_enough(acc[1], min_draws, stability_abs_tol, stability_rel_tol):
            break
    used = acc[0].n
    return {
        "observed_scores": observed.astype(np.float64), "null_center": center, "null_scale": scale,
        "null_positive_mass": acc[0].mean, "null_positive_units": acc[1].mean, "null_positive_fraction": acc[2].mean,
        "null_positive_units_sd": acc[1].sd(), "null_positive_units_stderr": acc[1].stderr(),
        "null_threshold": threshold, "null_tail_occupancy": acc[3].mean, "null_tail_occupancy_sd": acc[3].sd(),
        "null_tail_occupancy_stderr": acc[3].stderr(), "negative_support_size": support,
        "negative_fraction": float(support / max(int(ref_scores.size), 1)), "num_null_draws": used,
        "max_null_draws": draws, "adaptive_stop": bool(used < draws), "wild_bandwidth": bw,
        "wild_process": "bartlett_multiplier", "null_method": "dependent_wild_residual_bootstrap",
        "null_reference_mean_positive_consensus": float(np.mean(pos)),
        "null_reference_max_positive_consensus": float(np.max(pos)),
    }


def estimate_budget_nonnull_fraction(centered, observed_scores=None, dependence_lag_hint=None, return_details=False, **kw):
    """inference.py:988-1148: raw enriched fraction = clip(p_obs(s > t0) - E*[p(S* > t0)], 0, 1) plus the ESS metadata."""
    y = np.asarray(centered, dtype=np.float64)
    if y.ndim == 1:
        y = y[None, :]
    if y.ndim != 2:
        raise ValueError("`centered_matrix` must be one- or two-dimensional")
    n = y.shape[1]
    if n <= 0:
        raise ValueError("`centered_matrix` must contain at least one locus")
    meta = wild_bootstrap_score_null(y, observed_scores=observed_scores, dependence_lag_hint=dependence_lag_hint, **kw)
    return assemble_budget_details(meta, n, dependence_lag_hint, return_details)


def assemble_budget_details(meta, n, dependence_lag_hint, return_details, ess=None):
    obs = np.asarray(meta["observed_scores"], dtype=np.float64)
    center, scale = float(meta["null_center"]), float(meta["null_scale"])
    soft = float(max(scale, 1.0e-6))
    r = obs - center
    excess, deficit = np.clip(r, 0.0, None), np.clip(-r, 0.0, None)
    soft_counts = excess / soft
    L = ess_max_lag(n, dependence_lag_hint)
    n_eff, tau, used = effective_sample_size(soft_counts, L) if ess is None else ess
    tail_obs = float(np.mean(obs > float(meta["null_threshold"])))
    frac = float(np.clip(tail_obs - float(meta["null_tail_occupancy"]), 0.0, 1.0))
    if not (np.isfinite(frac) and np.isfinite(n_eff) and np.isfinite(tau)):
        raise ValueError("Budget initialization produced non-finite values")
    details = {
        "observed_positive_fraction": float(np.mean(excess > 0.0)),
        "observed_negative_fraction": float(np.mean(deficit > 0.0)),
        "null_positive_fraction": float(meta["null_positive_fraction"]),
        "observed_excess_mass": float(np.mean(excess)), "null_excess_mass": float(meta["null_positive_mass"]),
        "observed_excess_units": float(np.mean(soft_counts)), "null_excess_units": float(meta["null_positive_units"]),
        "null_excess_units_sd": float(meta["null_positive_units_sd"]),
        "null_excess_units_stderr": float(meta["null_positive_units_stderr"]),
        "null_threshold": float(meta["null_threshold"]), "observed_tail_occupancy": tail_obs,
        "null_tail_occupancy": float(meta["null_tail_occupancy"]),
        "null_tail_occupancy_sd": float(meta["null_tail_occupancy_sd"]),
        "null_tail_occupancy_stderr": float(meta["null_tail_occupancy_stderr"]),
        "null_center": center, "null_scale": scale, "nonnull_fraction": frac,
        "effective_count": float(frac * n_eff), "effective_total_count": float(n_eff),
        "autocorrelation_time": float(tau), "ess_max_lag": float(L), "ess_lags_used": float(used),
        "num_loci": float(n), "negative_support_size": float(meta["negative_support_size"]),
        "negative_fraction": float(meta["negative_fraction"]), "num_null_draws": float(meta["num_null_draws"]),
        "max_null_draws": float(meta["max_null_draws"]), "adaptive_stop": bool(meta["adaptive_stop"]),
        "wild_bandwidth": float(meta["wild_bandwidth"]), "wild_process": str(meta["wild_process"]),
        "null_method": str(meta["null_method"]),
        "null_reference_mean_positive_consensus": float(meta["null_reference_mean_positive_consensus"]),
        "null_reference_max_positive_consensus": float(meta["null_reference_max_positive_consensus"]),
    }
    return (frac, details) if return_details else frac


# ------------------------------------------------------------------ automatic gamma (rocco.py:751-789)
def resolve_chrom_gamma(chrom_scores, autocorrelation_time: float = 1.0, gamma=None):
    if gamma is not None:
        g = float(gamma)
        if not np.isfinite(g) or g < 0.0:
            raise ValueError("`--gamma` must be finite and non-negative")
        return g, None
    s = np.asarray(chrom_scores, dtype=np.float64)
    pos = s[s > 0.0]
    scale, count = (1.0, 0) if pos.size == 0 else (float(np.median(pos)), int(pos.size))
    tau = max(1.0, float(autocorrelation_time))
    run = int(np.ceil(tau))
    raw = 0.5 * float(run) * scale
    g = float(np.clip(raw, 0.5, 10.0))
    return g, {"method": "auto_score_autocorr", "autocorrelation_time": tau, "characteristic_run_length": run,
               "positive_score_median": scale, "positive_score_count": count, "gamma_raw": float(raw),
               "gamma_clipped": g, "gamma_clip_min": 0.5, "gamma_clip_max": 10.0}


# ------------------------------------------------------------------ narrowPeak summit offsets (rocco.py:840-872)
def narrowpeak_summit_offsets(track_starts, track_centers, track_mean, peak_starts, peak_ends) -> np.ndarray:
    """Per peak: arg-max of the (float32 -> float64) mean over the bins whose start lies in [start, end), NaN ignored,
    at least one finite value required; offset of that bin's centre clipped to the peak; else -1."""
    starts = np.asarray(track_starts, dtype=np.int64)
    centers = np.asarray(track_centers, dtype=np.int64)
    mean = np.asarray(np.asarray(track_mean, dtype=np.float32), dtype=np.float64)
    out = []
    for s, e in zip(peak_starts, peak_ends):
        off, length = -1, int(e) - int(s)
        if length > 0 and starts.size:
            lo, hi = int(np.searchsorted(starts, int(s), side="left")), int(np.searchsorted(starts, int(e), side="left"))
            if hi > lo and np.any(np.isfinite(mean[lo:hi])):
                k = int(np.nanargmax(mean[lo:hi]))
                off = int(np.clip(int(centers[lo + k]) - int(s), 0, max(length - 1, 0)))
        out.append(off)
    return np.array(out, dtype=np.int64)

"""GPU parity for SURVEY.md 8(f) ranks 1-2: dependent-wild-bootstrap budget null and automatic gamma.

Two kinds of check, because the reference's random streams are NumPy PCG64 and the product's are Philox:
  * REPLAY: the reference's own innovations (default_rng(seed + 104729 (draw + 1)), samples in order,
    inference.py:653-662) are handed to the CUDA path -> every number of the reference's details dict must come back
    (continuous statistics to 1e-6 relative -- they inherit the ~1e-7 score tolerance --, indicator means to 2 bins);
  * STATISTICAL: with its own Philox streams the multiplier field must have the reference's law (zero mean, unit
    variance, Bartlett autocorrelation, Gaussian marginals, independent samples) and the estimator must agree with
    the replayed one within Monte-Carlo error.
"""
import ctypes
import json
import os

import numpy as np
import pytest

from rocco_b200.synth import chrom_matrix_numpy

pytestmark = pytest.mark.gpu

CONTINUOUS_TOL = 1e-6
INDICATOR_KEYS = {"observed_positive_fraction", "observed_negative_fraction", "null_positive_fraction",
                  "observed_tail_occupancy", "null_tail_occupancy", "nonnull_fraction", "negative_fraction"}
COUNT_KEYS = {"negative_support_size", "effective_count"}          # scale with an indicator mean


@pytest.fixture(scope="module")
def budget_golden():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_budget_v1_11_0.npz"))


@pytest.fixture(scope="module")
def inf():
    from rocco_b200 import inference
    return inference


def reference_innovations(seed, draws, m, n, taps):
    out = np.empty((draws, m, n + taps - 1))
    for d in range(draws):
        rng = np.random.default_rng(int(seed) + 104729 * (d + 1))
        for i in range(m):
            out[d, i] = rng.standard_normal(n + taps - 1)
    return out


def wild_multiply(template, bandwidth, seed=0, draw=0, innovations=None):
    import torch
    from rocco_b200 import _lib
    lib = _lib.load()
    t = torch.as_tensor(np.ascontiguousarray(template, dtype=np.float64), device="cuda")
    out = torch.empty_like(t)
    inn = None if innovations is None else torch.as_tensor(np.ascontiguousarray(innovations, dtype=np.float64), device="cuda")
    st = lib.rocco_b200_wild_multiply_dev(ctypes.c_void_p(t.data_ptr()), t.shape[0], t.shape[1], int(bandwidth), int(seed), int(draw),
                                          None if inn is None else ctypes.c_void_p(inn.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                          ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(st, "wild multiply")
    torch.cuda.synchronize()
    return out.cpu().numpy()


# ------------------------------------------------------------------ replay of the reference's streams
@pytest.mark.parametrize("tag", ["ref_test", "synth_a", "synth_b"])
def test_budget_estimator_replays_reference(inf, budget_golden, tag):
    g = budget_golden
    kw = json.loads(str(g[f"{tag}_kwargs"]))
    want = json.loads(str(g[f"{tag}_meta"]))
    centered, scores = g[f"{tag}_centered"], g[f"{tag}_scores"]
    m, n = centered.shape
    bw = inf._resolve_budget_bootstrap_bandwidth(n, kw.get("dependence_lag_hint"))
    assert bw == int(want["wild_bandwidth"])
    inn = reference_innovations(kw.get("random_seed", 0), kw["num_null_draws"], m, n, 2 * bw + 1)
    frac, meta = inf.estimate_budget_nonnull_fraction_from_wild_bootstrap_null(
        centered, observed_scores=scores, return_details=True, innovations=inn, **kw)
    assert set(want) <= set(meta)
    for k, v in want.items():
        if isinstance(v, (str, bool)):
            assert meta[k] == v, k
        elif k in INDICATOR_KEYS:
            assert abs(meta[k] - v) <= 2.0 / n + 1e-12, (k, meta[k], v)
        elif k in COUNT_KEYS:
            assert abs(meta[k] - v) <= 2.0 + 1e-6 * abs(v), (k, meta[k], v)
        else:
            assert abs(meta[k] - v) <= CONTINUOUS_TOL * max(1.0, abs(v)), (k, meta[k], v)
    assert abs(frac - float(g[f"{tag}_fraction"])) <= 2.0 / n


def test_gamma_and_ess_match_reference(inf, budget_golden):
    from rocco_b200.rocco import _resolve_chrom_gamma
    g = budget_golden
    for key, lag, vals in (("ess_out", 404, g["ess_values"]), ("ess_out_short", 64, g["ess_values"][:40])):
        n_eff, tau, used = inf._estimate_effective_sample_size(vals, lag)
        assert abs(n_eff - g[key][0]) <= 1e-9 * g[key][0] and abs(tau - g[key][1]) <= 1e-9 * g[key][1] and used == int(g[key][2])
    assert inf._estimate_effective_sample_size(np.ones(3), 10) == (3.0, 1.0, 0)
    assert inf._estimate_effective_sample_size(np.full(100, 2.5), 10) == (100.0, 1.0, 0)
    with pytest.raises(ValueError):
        inf._estimate_effective_sample_size(np.ones((2, 2)), 4)
    for tag in ("ref_test", "synth_a", "synth_b"):
        meta = json.loads(str(g[f"{tag}_meta"]))
        want = json.loads(str(g[f"{tag}_gamma_meta"]))
        gam, gm = _resolve_chrom_gamma(tag, {"gamma": None}, g[f"{tag}_scores"], meta)      # reference meta: median on the GPU
        assert gam == float(g[f"{tag}_gamma"])
        for k, v in want.items():
            assert gm[k] == v, (k, gm[k], v)
    assert _resolve_chrom_gamma("c", {"gamma": 2.5}, np.zeros(4), {}) == (2.5, None)
    assert _resolve_chrom_gamma("c", {"gamma": None}, -np.ones(7), {})[0] == 0.5            # no positive score: scale 1
    with pytest.raises(ValueError):
        _resolve_chrom_gamma("c", {"gamma": -1.0}, np.zeros(4), {})
    for n in (1, 2, 9, 512, 3000, 934200):
        for h in (None, 16, 25, 101, 1000):
            row = [r for r in g["bandwidth_rules"] if r[0] == n and r[1] == (-1 if h is None else h)][0]
            assert inf._resolve_budget_bootstrap_bandwidth(n, h) == row[2] and inf._resolve_budget_ess_max_lag(n, h) == row[3]
    assert np.array_equal(inf._build_budget_bootstrap_kernel(101), g["kernel_b101"])


def test_wild_field_replay_matches_oracle_weights():
    """FIR + centring + scaling against the oracle's restatement of inference.py:544-570, across tile edges."""
    from oracle import budget as ob
    rng = np.random.default_rng(3)
    for n, bw in ((1, 1), (2, 1), (9, 8), (4095, 16), (4096, 101), (4097, 8), (20000, 171)):
        taps = ob.bartlett_kernel(bw)
        m = 3
        inn = rng.standard_normal((m, n + 2 * bw))
        tmpl = rng.standard_normal((m, n))
        got = wild_multiply(tmpl, bw, innovations=inn)
        if n == 1:
            want = tmpl.copy()
        else:
            want = np.stack([tmpl[i] * ob.dependent_wild_weights(inn[i], taps) for i in range(m)])
        assert np.max(np.abs(got - want)) <= 1e-12 * max(1.0, np.max(np.abs(want))), (n, bw)


def test_philox_field_has_the_reference_law():
    from oracle import budget as ob
    n, bw, m = 1 << 20, 16, 4
    w = wild_multiply(np.ones((m, n)), bw, seed=1234, draw=0)
    assert np.array_equal(w, wild_multiply(np.ones((m, n)), bw, seed=1234, draw=0))          # reproducible
    assert not np.array_equal(w[0], w[1])
    assert np.max(np.abs(w.mean(axis=1))) < 1e-12 and np.max(np.abs(w.std(axis=1) - 1.0)) < 1e-12
    taps = ob.bartlett_kernel(bw)
    for lag in (1, 4, 16, 32, 33, 64):
        want = float(np.dot(taps[: taps.size - lag], taps[lag:])) if lag < taps.size else 0.0
        got = float(np.mean(w[:, :-lag] * w[:, lag:]))
        assert abs(got - want) < 0.02, (lag, got, want)
    z = w[:, :: 4 * bw].ravel()                                                               # ~independent subsample
    assert abs(np.mean(z ** 3)) < 0.1 and abs(np.mean(z ** 4) - 3.0) < 0.2
    assert abs(np.corrcoef(w[0], w[1])[0, 1]) < 0.02                                          # samples are independent
    w2 = wild_multiply(np.ones((m, n)), bw, seed=1234, draw=1)
    assert abs(np.corrcoef(w[0], w2[0])[0, 1]) < 0.02                                         # so are draws
    w3 = wild_multiply(np.ones((m, n)), bw, seed=99, draw=0)
    assert abs(np.corrcoef(w[0], w3[0])[0, 1]) < 0.02                                         # and seeds


def test_philox_estimator_agrees_with_replayed_streams_in_distribution(inf, budget_golden):
    g = budget_golden
    centered, scores = g["synth_a_centered"], g["synth_a_scores"]
    common = dict(observed_scores=scores, dependence_lag_hint=101, prior_df=6.0, num_null_draws=24, min_null_draws=24,
                  return_details=True)
    runs = [inf.estimate_budget_nonnull_fraction_from_wild_bootstrap_null(centered, random_seed=s, **common)[1] for s in (0, 1, 2)]
    want = json.loads(str(g["synth_a_meta"]))                 # 5 draws of the reference's own streams
    for meta in runs:
        assert meta["num_null_draws"] == 24.0 and meta["null_center"] == runs[0]["null_center"]
        for k, sd_key in (("null_excess_units", "null_excess_units_sd"), ("null_tail_occupancy", "null_tail_occupancy_sd")):
            sd = max(meta[sd_key], want[sd_key])
            assert abs(meta[k] - want[k]) <= 4.0 * sd * np.sqrt(1.0 / 24 + 1.0 / 5), (k, meta[k], want[k], sd)
    assert runs[0]["null_excess_units"] != runs[1]["null_excess_units"]
    again = inf.estimate_budget_nonnull_fraction_from_wild_bootstrap_null(centered, random_seed=0, **common)[1]
    assert again == runs[0]                                   # deterministic for a given seed


def test_reference_test_case_assertions_hold_with_philox(inf):
    """tests/test_rocco.py:462-500 of the reference, run through the CUDA path with its default generator."""
    import rocco_b200
    x = np.arange(512, dtype=np.float64)
    p1 = 6.0 * np.exp(-0.5 * ((x - 120.0) / 15.0) ** 2)
    p2 = 5.5 * np.exp(-0.5 * ((x - 320.0) / 15.0) ** 2)
    mat = np.vstack([0.25 + p1 + p2 + 0.05 * np.sin(x / 13.0), 0.20 + 0.95 * p1 + 1.05 * p2 + 0.04 * np.cos(x / 15.0),
                     0.22 + 1.1 * p1 + 0.9 * p2 + 0.05 * np.sin(x / 17.0)])
    scores, details = rocco_b200.score_loci_wls(mat, return_details=True)
    fraction, meta = rocco_b200.estimate_budget_nonnull_fraction_from_empirical_null(
        details["centered_matrix"], observed_scores=scores, dependence_lag_hint=16, num_null_draws=6, return_details=True)
    assert 0.0 < fraction <= 1.0 and np.isclose(fraction, meta["nonnull_fraction"])
    assert 0.0 <= meta["observed_positive_fraction"] <= 1.0 and 0.0 <= meta["null_positive_fraction"] <= 1.0
    assert meta["observed_excess_mass"] > meta["null_excess_mass"] > 0.0
    assert meta["observed_excess_units"] > meta["null_excess_units"] > 0.0
    assert meta["effective_count"] > 0.0 and 1.0 <= meta["effective_total_count"] <= meta["num_loci"]
    assert meta["autocorrelation_time"] >= 1.0 and meta["ess_max_lag"] == 64.0
    assert meta["null_method"] == "dependent_wild_residual_bootstrap"
    assert meta["num_null_draws"] == 6.0 and meta["max_null_draws"] == 6.0 and not meta["adaptive_stop"]
    assert meta["wild_bandwidth"] >= 8.0 and meta["null_excess_units_sd"] > 0.0
    assert meta["null_reference_mean_positive_consensus"] >= 0.0
    assert meta["negative_support_size"] > 0.0 and 0.0 < meta["negative_fraction"] <= 1.0


def test_budget_errors_and_edge_shapes(inf):
    with pytest.raises(ValueError):
        inf.estimate_budget_nonnull_fraction_from_wild_bootstrap_null(np.zeros((2, 3, 4)))
    with pytest.raises(ValueError):
        inf.estimate_budget_nonnull_fraction_from_wild_bootstrap_null(np.zeros((2, 0)))
    x = chrom_matrix_numpy(3, 400, seed=2)
    with pytest.raises(ValueError):
        inf.estimate_budget_nonnull_fraction_from_wild_bootstrap_null(x, observed_scores=np.zeros(5))
    # a one-dimensional centred track is one sample (inference.py:1041-1042)
    c = np.random.default_rng(0).standard_normal(600) * 0.3
    f1 = inf.estimate_budget_nonnull_fraction_from_wild_bootstrap_null(c, num_null_draws=4, random_seed=5)
    f2 = inf.estimate_budget_nonnull_fraction_from_wild_bootstrap_null(c[None, :], num_null_draws=4, random_seed=5)
    assert f1 == f2 and 0.0 <= f1 <= 1.0


def test_budget_chr21_size_runs_and_is_sane(inf):
    """BASELINE config-1 size (934,200 bins x 100 samples would need the 750 MB centred matrix on the host twice; 20 samples
    keep the test's host side small): invariants only, plus a timing line for the log."""
    import time
    import rocco_b200
    x = chrom_matrix_numpy(20, 934200, seed=21)
    scores, det = rocco_b200.score_loci_wls(x, prior_df=6.0, return_details=True)
    t0 = time.perf_counter()
    frac, meta = inf.estimate_budget_nonnull_fraction_from_wild_bootstrap_null(
        det["centered_matrix"], observed_scores=scores, prior_df=6.0, dependence_lag_hint=int(det["local_baseline_window"]),
        num_null_draws=25, return_details=True)
    dt = time.perf_counter() - t0
    print(f"\nbudget null 20 x 934200, {int(meta['num_null_draws'])} draws: {dt * 1e3:.0f} ms")
    assert 0.0 < frac < 0.2                              # ~2 % of the synthetic bins carry peaks
    assert 8 <= meta["num_null_draws"] <= 25 and meta["wild_bandwidth"] == 101.0
    assert 0.0 <= meta["null_tail_occupancy"] < meta["observed_tail_occupancy"] < 0.2
    assert 1.0 <= meta["autocorrelation_time"] < 100.0
    assert meta["positive_score_count"] == int(np.sum(scores > 0))
    assert meta["positive_score_median"] == float(np.median(scores[scores > 0]))


def test_narrowpeak_summit_offsets_match_reference(budget_golden, tmp_path):
    """rocco.py:809-872 through the CUDA path: the reference's own output file must be reproduced byte for byte
    (NaN / +-inf means, tie plateaus, peaks off the track, zero-length peaks, a chromosome without a track)."""
    from rocco_b200.rocco import _cpy_narrowpeak_summit_track, _write_narrowpeak_summit_offsets, narrowpeak_summit_offsets
    g = budget_golden
    cache = {c: {"summit_track_file": _cpy_narrowpeak_summit_track(c, g[f"summit_{c}_intervals"], g[f"summit_{c}_mean"])}
             for c in ("chrA", "chrB")}
    peak_file = tmp_path / "peaks.bed"
    peak_file.write_text(str(g["summit_peaks_text"]))
    out = _write_narrowpeak_summit_offsets(str(peak_file), cache, str(tmp_path / "offsets.tsv"))
    assert open(out).read() == str(g["summit_offsets_text"])
    assert _cpy_narrowpeak_summit_track("c", np.array([5]), np.array([1.0])) is None
    assert narrowpeak_summit_offsets([], [], [], [1, 2], [5, 2]).tolist() == [-1, -1]
    for f in cache.values():
        os.remove(f["summit_track_file"])

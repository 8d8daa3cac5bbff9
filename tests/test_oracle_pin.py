"""Pins the oracle (CPU, no GPU needed).

1. oracle port (oracle/c/*.c + oracle/oracle.py) == golden vectors produced by the REAL reference
   (tests/golden/make_golden.py, ROCCO v1.11.0) -- bit for bit.
2. oracle port == oracle/_ref (the reference's own extensions compiled here) on fresh seeded inputs.
3. The known answers the reference's own tests hold (reference tests/test_rocco.py:234-246,
   331-345, 397-437, 895-896) and its BED fixtures (combined_ref.bed).
"""
import itertools
import os

import numpy as np
import pytest

from rocco_b200.synth import chrom_matrix_numpy


def _eq(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    assert np.array_equal(a, b), f"max abs diff {np.max(np.abs(a.astype(float) - b.astype(float)))}"


KINDS = ["port", "reference"]


def _kinds(oracle):
    return [k for k in KINDS if k == "port" or oracle.reference_available()]


# ---------------------------------------------------------------- golden: scoring
@pytest.mark.parametrize("tag,kw", [("a", {}),
                                    ("b", dict(prior_df=6.0, lower_bound_z=0.5, precision_floor_ratio=0.05)),
                                    ("c", dict(min_effect=0.25))])
def test_score_loci_wls_matches_golden(oracle, golden, tag, kw):
    for kind in _kinds(oracle):
        sc, det = oracle.score_loci_wls(golden["score_x"], return_details=True, kind=kind, **kw)
        _eq(sc, golden[f"score_{tag}_scores"])
        for k in ("mean", "raw_variance", "prior_variance", "moderated_variance", "standard_error", "z_scores"):
            _eq(det[k], golden[f"score_{tag}_{k}"])
        if tag == "a":
            _eq(det["centered_matrix"], golden["score_a_centered"])
            meta = golden["score_a_meta"]
            assert det["local_baseline_window"] == int(meta[0]) == 101
            assert det["local_baseline_lambda"] == meta[1]
            assert det["prior_spatial_window"] == int(meta[2]) == 31
            assert det["degrees_of_freedom"][0] == meta[3]


def test_score_loci_wls_float32_input_matches_golden(oracle, golden):
    _eq(oracle.score_loci_wls(golden["score_x"].astype(np.float32)), golden["score_f32_scores"])


@pytest.mark.parametrize("tag", ["n2", "n3", "n4", "n24", "n25", "n130"])
def test_small_n_branches_match_golden(oracle, golden, tag):
    for kind in _kinds(oracle):
        sc, det = oracle.score_loci_wls(golden[f"small_{tag}_x"], lower_bound_z=0.0, return_details=True, kind=kind)
        _eq(sc, golden[f"small_{tag}_scores"])
        _eq(det["mean"], golden[f"small_{tag}_mean"])
        _eq(det["standard_error"], golden[f"small_{tag}_se"])


def test_stage_baseline_and_wls_match_golden(oracle, golden):
    for kind in _kinds(oracle):
        _eq(oracle.estimate_local_background_matrix(golden["base_y"], kind=kind)[0], golden["base_out"])
        _eq(oracle.crossfit_whittaker_baseline_1d(golden["base129_y"], block_size=41, kind=kind), golden["base129_out"])
        sc, det = oracle.score_centered_wls_matrix(golden["wls_centered"], prior_df=6.0, spatial_window=31, kind=kind)
        _eq(sc, golden["wls_scores"])
        for k in ("mean", "raw_variance", "prior_variance", "moderated_variance", "standard_error"):
            _eq(det[k], golden[f"wls_{k}"])


# ---------------------------------------------------------------- golden: DP + search
@pytest.mark.parametrize("tag", ["g1", "g7", "g0"])
def test_solve_chrom_exact_matches_golden(oracle, golden, tag):
    budget, gamma, obj, pen, cnt, lam = golden[f"dp_{tag}_meta"]
    for kind in _kinds(oracle):
        sol, o, det = oracle.solve_chrom_exact(golden["score_a_scores"], budget=budget, gamma=gamma,
                                               return_details=True, kind=kind)
        _eq(sol, golden[f"dp_{tag}_mask"])
        assert (o, det["penalized_objective"], det["selected_count"], det["selection_penalty"]) == (obj, pen, int(cnt), lam)


@pytest.mark.parametrize("tag", ["const", "ints"])
def test_tie_break_cases_match_golden(oracle, golden, tag):
    arr = golden[f"dp_tie_{tag}_scores"]
    for k in range(3):
        b, gm, obj, pen, cnt, lam = golden[f"dp_tie_{tag}_{k}_meta"]
        sol, o, det = oracle.solve_chrom_exact(arr, budget=None if b < 0 else b, gamma=gm, return_details=True)
        _eq(sol, golden[f"dp_tie_{tag}_{k}_mask"])
        assert (o, det["penalized_objective"], det["selected_count"], det["selection_penalty"]) == (obj, pen, int(cnt), lam)


def _bruteforce(scores, costs, pen):
    best = (None, -np.inf, None)
    for bits in itertools.product([0, 1], repeat=len(scores)):
        z = np.asarray(bits, dtype=np.uint8)
        v = scores @ z - np.sum(costs * np.abs(np.diff(z))) - pen * np.sum(z)
        if v > best[1] or (np.isclose(v, best[1]) and np.sum(z) < best[2]):
            best = (z, float(v), int(np.sum(z)))
    return best


def test_exact_dp_matches_bruteforce_and_golden(oracle, golden):
    """reference tests/test_rocco.py:397-415"""
    rng = np.random.default_rng(7)
    scores, costs = rng.normal(size=9), rng.uniform(0.2, 1.3, size=8)
    _eq(scores, golden["dp_bf_scores"])
    for k, pen in enumerate((-0.5, 0.0, 0.6, 1.4)):
        for kind in _kinds(oracle):
            sol, val, cnt = oracle.solve_penalized_chain(scores, costs, pen, kind=kind)
            bsol, bval, bcnt = _bruteforce(scores, costs, pen)
            _eq(sol, bsol)
            assert np.isclose(val, bval) and cnt == bcnt
            _eq(sol, golden[f"dp_bf_{k}_mask"])
            assert (pen, val, cnt) == tuple(golden[f"dp_bf_{k}_meta"])


def test_solve_chrom_exact_respects_budget(oracle, golden):
    """reference tests/test_rocco.py:418-437; SURVEY appendix C: mask [0,0,0,0,1,1,0,0], lambda 1.05, objective -3.8"""
    s = np.array([0.5, 1.5, 1.4, -0.2, 3.0, 2.8, -0.1, 0.1])
    sol, obj, det = oracle.solve_chrom_exact(s, budget=0.375, gamma=1.0, return_details=True)
    assert sol.dtype == np.uint8 and sol.tolist() == [0, 0, 0, 0, 1, 1, 0, 0]
    assert np.sum(sol) <= 3 and det["selected_fraction"] <= 0.375
    assert np.isclose(obj, oracle.objective_value(sol, s, oracle.build_switch_costs(s, 1.0)))
    assert np.isclose(obj, -3.8) and np.isclose(det["selection_penalty"], 1.05)
    _eq(sol, golden["dp_s8_mask"])
    assert (obj, det["penalized_objective"], det["selected_count"], det["selection_penalty"]) == tuple(golden["dp_s8_meta"])


# ---------------------------------------------------------------- reference known answers (scoring)
def test_known_answer_log_scale(oracle):
    """reference tests/test_rocco.py:234-246"""
    sc, det = oracle.score_loci_wls(np.array([[1.0, 15.0]]), lower_bound_z=0.0, return_details=True)
    assert det["input_scale"] == "log2p1"
    assert np.allclose(det["mean"], [-1.5, 1.5])
    assert np.allclose(det["z_scores"], [-0.67449076, 0.67449076])
    assert np.allclose(sc, [-0.67449076, 0.67449076])


def test_known_answer_tied_large_matrix(oracle):
    """reference tests/test_rocco.py:331-345 (all-ties path through sort / PAVA)"""
    sc, det = oracle.score_centered_wls_matrix(np.zeros((3, 250000)), lower_bound_z=1.0, prior_df=5.0)
    assert sc.shape == (250000,)
    assert np.allclose(det["mean"], 0.0) and np.allclose(det["z_scores"], 0.0) and np.allclose(sc, -1.0)
    assert np.all(det["standard_error"] > 0.0)


def test_known_answer_min_effect_and_precision_floor(oracle):
    """reference tests/test_rocco.py:249-285"""
    sc, det = oracle.score_loci_wls(np.array([[1.0, 15.0]]), min_effect=0.5, return_details=True)
    assert np.isclose(det["min_effect"], 0.5) and np.all(sc < det["z_scores"])
    c = np.array([[0.05, 1.0, 1.0, 0.05], [0.04, 1.0, 1.0, 0.04], [0.06, 1.0, 1.0, 0.06]])
    lo_s, lo_d = oracle.score_centered_wls_matrix(c, prior_df=6.0, precision_floor_ratio=0.0)
    hi_s, hi_d = oracle.score_centered_wls_matrix(c, prior_df=6.0, precision_floor_ratio=0.25)
    assert np.all(hi_d["standard_error"] >= lo_d["standard_error"]) and np.all(hi_s <= lo_s)


def test_known_answer_low_memory_dtype(oracle):
    """reference tests/test_rocco.py:288-297"""
    sc, det = oracle.score_loci_wls(np.array([[1.0, 3.0, 7.0], [1.2, 2.8, 6.5]]), low_memory=True, return_details=True)
    assert sc.dtype == np.float64 and det["centered_matrix"].dtype == np.float32


# ---------------------------------------------------------------- port == compiled reference on fresh inputs
@pytest.mark.parametrize("m,n,seed", [(3, 2000, 1), (8, 5000, 2), (2, 40, 3), (5, 26, 4)])
def test_port_equals_compiled_reference_scoring(oracle, m, n, seed):
    if not oracle.reference_available():
        pytest.skip("oracle/_ref not built")
    x = chrom_matrix_numpy(m, n, seed=seed)
    a, da = oracle.score_loci_wls(x, prior_df=6.0, return_details=True, kind="port")
    b, db = oracle.score_loci_wls(x, prior_df=6.0, return_details=True, kind="reference")
    _eq(a, b)
    for k in ("mean", "raw_variance", "prior_variance", "moderated_variance", "standard_error", "centered_matrix"):
        _eq(da[k], db[k])


@pytest.mark.parametrize("n,gamma,budget,seed", [(4000, 1.0, 0.02, 0), (4000, 6.86, 0.05, 1), (777, 0.25, 0.3, 2)])
def test_port_equals_compiled_reference_search(oracle, n, gamma, budget, seed):
    if not oracle.reference_available():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(seed)
    s = rng.normal(size=n) + 3.0 * (rng.random(n) < 0.03)
    ta, tb = [], []
    c = oracle.build_switch_costs(s, gamma)
    a = oracle.calibrate_selection_penalty(s, c, int(np.floor(n * budget)), kind="port", trace=ta)
    b = oracle.calibrate_selection_penalty(s, c, int(np.floor(n * budget)), kind="reference", trace=tb)
    assert ta == tb and len(ta) == 62
    assert a[0] == b[0] and a[2] == b[2] and a[3] == b[3]
    _eq(a[1], b[1])
    rc = rng.uniform(0.0, 2.0, size=n - 1)
    for pen in (-1.0, 0.0, 0.37, 2.0):
        pa, pb = (oracle.solve_penalized_chain(s, rc, pen, kind=k) for k in ("port", "reference"))
        _eq(pa[0], pb[0])
        assert pa[1:] == pb[1:]


# ---------------------------------------------------------------- column statistics and BED
def test_column_statistics_match_golden(oracle, golden):
    x = golden["col_x"]
    _eq(oracle.score_central_tendency_chrom(x), golden["col_median"])
    _eq(oracle.score_central_tendency_chrom(x, method="quantile", quantile=0.75), golden["col_q75"])
    _eq(oracle.score_central_tendency_chrom(x, method="quantile", quantile=0.25, power=0.5), golden["col_q25_pow"])
    _eq(oracle.score_central_tendency_chrom(x, method="tmean", tprop=0.1), golden["col_tmean"])
    _eq(oracle.score_central_tendency_chrom(x, method="mean"), golden["col_mean"])
    _eq(oracle.score_dispersion_chrom(x, method="mad"), golden["col_mad"])
    _eq(oracle.score_dispersion_chrom(x, method="iqr"), golden["col_iqr"])
    _eq(oracle.score_dispersion_chrom(x, method="std"), golden["col_std"])
    x10 = golden["col10_x"]
    _eq(oracle.score_central_tendency_chrom(x10), golden["col10_median"])
    _eq(oracle.score_central_tendency_chrom(x10, method="quantile", quantile=0.75), golden["col10_q75"])
    _eq(oracle.score_dispersion_chrom(x10, method="mad"), golden["col10_mad"])
    _eq(oracle.score_dispersion_chrom(x10, method="iqr", rng=(10, 90)), golden["col10_iqr"])
    # reference tests/test_rocco.py:895-896 (bigWig path uses the column median)
    bw = oracle.score_central_tendency_chrom(np.array([[0.0, 2.0, 1.0, 0.0], [0.0, 3.0, 2.0, 0.0]]))
    assert bw.tolist() == [0.0, 2.5, 1.5, 0.0]
    _eq(bw, golden["col_bw"])


def test_bed_text_matches_golden(oracle, golden, tmp_path):
    iv = golden["bed_iv"]
    for tag, ml in (("all", None), ("min150", 150)):
        f = oracle.chrom_solution_to_bed("chr21", iv, golden["dp_g1_mask"], ID="gold", min_length_bp=ml, out_dir=tmp_path)
        assert open(f, "rb").read() == golden[f"bed_{tag}_text"].tobytes()
    toy = np.array([0, 1, 1, 0, 0, 1, 0, 1, 1, 1], dtype=np.uint8)
    f = oracle.chrom_solution_to_bed("chrT", np.arange(0, 500, 50), toy, out_dir=tmp_path)
    text = open(f, "rb").read()
    assert text == golden["bed_toy_text"].tobytes()
    assert text.decode().splitlines()[-1] == "chrT\t350\t450"      # last bin never emitted


def test_combine_reproduces_reference_combined_bed(oracle, bed_fixtures, tmp_path):
    """reference tests/test_rocco.py:216-231 asserts Jaccard > 0.99; identical records are required here."""
    files = []
    for name in ("ref_chr19", "ref_chr21", "ref_chrX"):
        p = tmp_path / f"{name}.bed"
        p.write_bytes(bed_fixtures[name].tobytes())
        files.append(str(p))
    out = oracle.combine_chrom_results(files, str(tmp_path / "combined.bed"))
    assert open(out, "rb").read() == bed_fixtures["combined_ref"].tobytes()
    # budget sanity: selected 50-bp bins ~ hg_params budgets
    for name, bins in (("ref_chr21", 16675), ("ref_chr19", 52689), ("ref_chrX", 46805)):
        recs = oracle.read_bed_records(str(tmp_path / f"{name}.bed"))
        assert sum(e - s for _, s, e in recs) // 50 == bins


# ------------------------------------------------------------------------------------------------
# dependent-wild-bootstrap budget null + automatic gamma (oracle/budget.py)
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def budget_golden():
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_budget_v1_11_0.npz"))


def _close(a, b, tol):
    return abs(float(a) - float(b)) <= tol * max(1.0, abs(float(b)))


def test_budget_building_blocks_match_reference(budget_golden):
    from oracle import budget as ob
    g = budget_golden
    for b in (8, 101):
        assert np.array_equal(ob.bartlett_kernel(b), g[f"kernel_b{b}"])
    for n, h, bw, lag in g["bandwidth_rules"]:
        hint = None if h < 0 else int(h)
        assert ob.bootstrap_bandwidth(int(n), hint) == bw and ob.ess_max_lag(int(n), hint) == lag
    taps = ob.bartlett_kernel(16)
    w = ob.dependent_wild_weights(np.random.default_rng(77).standard_normal(4000 + taps.size - 1), taps)
    assert np.max(np.abs(w - g["weights_n4000_b16"])) < 1e-12          # FFT vs direct correlation: rounding only
    for key, vals, lag in (("ess_out", g["ess_values"], 404), ("ess_out_short", g["ess_values"][:40], 64)):
        n_eff, tau, used = ob.effective_sample_size(vals, lag)
        assert _close(n_eff, g[key][0], 1e-10) and _close(tau, g[key][1], 1e-10) and used == int(g[key][2])


@pytest.mark.parametrize("tag", ["ref_test", "synth_a", "synth_b"])
def test_budget_estimator_matches_reference(budget_golden, tag):
    """Same PCG64 streams as the reference => every entry of the details dict is reproduced."""
    import json
    from oracle import budget as ob
    g = budget_golden
    kw = json.loads(str(g[f"{tag}_kwargs"]))
    want = json.loads(str(g[f"{tag}_meta"]))
    frac, meta = ob.estimate_budget_nonnull_fraction(g[f"{tag}_centered"], observed_scores=g[f"{tag}_scores"],
                                                     return_details=True, **kw)
    assert set(meta) == set(want)
    for k, v in want.items():
        if isinstance(v, (str, bool)):
            assert meta[k] == v, k
        else:
            assert _close(meta[k], v, 1e-9), (k, meta[k], v)
    assert _close(frac, float(g[f"{tag}_fraction"]), 1e-9)
    gam, gmeta = ob.resolve_chrom_gamma(g[f"{tag}_scores"], meta["autocorrelation_time"])
    wantg = json.loads(str(g[f"{tag}_gamma_meta"]))
    assert _close(gam, float(g[f"{tag}_gamma"]), 1e-12)
    for k, v in wantg.items():
        assert (gmeta[k] == v) if isinstance(v, str) else _close(gmeta[k], v, 1e-10), k


def _summit_case(g):
    peaks = [ln.split("\t") for ln in str(g["summit_peaks_text"]).strip().split("\n")]
    want = [ln.split("\t") for ln in str(g["summit_offsets_text"]).strip().split("\n")]
    assert [f"{c}_{s}_{e}" for c, s, e in peaks] == [w[0] for w in want]
    tracks = {}
    for chrom in ("chrA", "chrB"):
        iv, mean = g[f"summit_{chrom}_intervals"], g[f"summit_{chrom}_mean"]
        usable = min(iv.shape[0] - 1, mean.shape[0])
        tracks[chrom] = (iv[:usable], (iv[:usable] + iv[1:usable + 1]) // 2, mean[:usable].astype(np.float32))
    return peaks, np.array([int(w[1]) for w in want]), tracks


def test_narrowpeak_summit_offsets_match_reference(budget_golden):
    from oracle import budget as ob
    peaks, want, tracks = _summit_case(budget_golden)
    got = np.full(len(peaks), -1, dtype=np.int64)
    for chrom, (ts, tc, tm) in tracks.items():
        idx = [k for k, p in enumerate(peaks) if p[0] == chrom]
        got[idx] = ob.narrowpeak_summit_offsets(ts, tc, tm, [int(peaks[k][1]) for k in idx], [int(peaks[k][2]) for k in idx])
    assert np.array_equal(got, want)
    assert (want == -1).sum() > 3 and (want >= 0).sum() > 400

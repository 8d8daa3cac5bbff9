"""Size-independent properties at BASELINE.json's full per-chromosome sizes (chr1 @ 50 bp: 4,979,129 bins; chr1 @ 20 bp:
12,447,822 bins) where the CPU oracle would take minutes: round trips, monotonicity, linearity, invariances."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CHR1_50 = 4_979_129
CHR1_20 = 12_447_822


def _scores(n, seed):
    rng = np.random.default_rng(seed)
    return rng.normal(size=n) - 0.8 + 4.0 * (rng.random(n) < 0.03) * rng.random(n)


@pytest.mark.parametrize("n,budget,gamma", [(CHR1_50, 0.03, 1.0), (CHR1_20, 0.03, 6.86)])
def test_budget_solve_properties_full_chr1(n, budget, gamma):
    import torch
    from rocco_b200 import pipeline
    s = _scores(n, 1)
    d, offs, lens = pipeline.pack_scores([s])
    d_masks, res = pipeline.solve_packed(d, offs, lens, [budget], [gamma])
    r = res[0]
    target = pipeline.target_count_for_budget(n, budget)
    mask = d_masks[:n].cpu().numpy()
    assert r["dp_passes"] == 62
    assert int(mask.sum()) == r["selected_count"] <= target
    # the objective the kernel reports is the objective of the mask it emitted (dp.py:16-34)
    tv = gamma * float(np.abs(np.diff(mask.astype(np.int8))).sum())
    want_obj = -float(s @ mask) + tv
    assert abs(r["objective"] - want_obj) <= 1e-9 * max(1.0, abs(want_obj))
    assert abs(r["penalized_objective"] - (float(s @ mask) - r["selection_penalty"] * r["selected_count"] - tv)) <= 1e-6 * abs(want_obj)
    # just below the returned multiplier the count exceeds the target (the search returns the feasible side of a breakpoint)
    lam = r["selection_penalty"]
    counts, pen, obj = pipeline.sweep_multipliers(d[:n], gamma, [lam - 1e-6, lam, lam + 1e-6, lam + 0.5])
    assert counts[1] == r["selected_count"] and counts[0] > target >= counts[1] >= counts[2] >= counts[3]
    # optimality at fixed lambda: the penalized value of the DP mask is >= that of a few perturbed masks
    rng = np.random.default_rng(0)
    a = s - lam
    base = float(a @ mask) - tv
    for _ in range(20):
        alt = mask.copy()
        i = int(rng.integers(0, n - 60)); w = int(rng.integers(1, 50))
        alt[i:i + w] ^= 1
        val = float(a @ alt) - gamma * float(np.abs(np.diff(alt.astype(np.int8))).sum())
        assert val <= base + 1e-7


def test_sweep_is_monotone_and_convex_full_chr1():
    from rocco_b200 import pipeline
    s = _scores(CHR1_50, 2)
    lams = np.linspace(np.quantile(s[:200000], 0.5), np.quantile(s[:200000], 0.999), 256)
    counts, pen, obj = pipeline.sweep_multipliers(s, 1.0, lams)
    assert np.all(np.diff(counts) <= 0)
    assert np.all(np.diff(pen) <= 1e-6)                            # max_z of an affine-in-lambda family: non-increasing
    slopes = np.diff(pen) / np.diff(lams)
    assert np.all(np.diff(slopes) >= -1e-3 * np.abs(slopes[:-1]) - 1e-3)   # ... and convex; slope = -count
    assert np.all(-slopes <= counts[:-1] + 1e-6 * counts[:-1] + 1) and np.all(-slopes >= counts[1:] - 1e-6 * counts[1:] - 1)


def test_mask_to_intervals_round_trip_full_chr1(tmp_path, monkeypatch):
    import rocco_b200
    from rocco_b200.rocco import solution_runs
    rng = np.random.default_rng(3)
    mask = (rng.random(CHR1_20) < 0.2).astype(np.uint8)
    first, last = solution_runs(mask)
    rebuilt = np.zeros(CHR1_20, dtype=np.uint8)
    delta = np.zeros(CHR1_20 + 1, dtype=np.int32)
    np.add.at(delta, first, 1)
    np.add.at(delta, last, -1)
    rebuilt[:] = np.cumsum(delta[:-1]) > 0
    assert np.array_equal(rebuilt[:-1], mask[:-1]) and rebuilt[-1] == 0         # the last bin is never emitted
    assert np.all(first[1:] > last[:-1])                                          # maximal, disjoint, sorted runs


def test_baseline_linearity_and_constants_full_chr1():
    from rocco_b200 import _baseline
    from rocco_b200.inference import _consenrich_whittaker_lambda
    lam = _consenrich_whittaker_lambda(101)
    rng = np.random.default_rng(4)
    n = CHR1_50
    y = np.vstack([rng.normal(size=n).cumsum() * 1e-3 + rng.normal(size=n), rng.normal(size=n)])
    b = _baseline.crossfit_whittaker_baseline(y, lam)
    comb = _baseline.crossfit_whittaker_baseline(2.0 * y[0] - 0.5 * y[1], lam)
    assert np.max(np.abs(comb - (2.0 * b[0] - 0.5 * b[1]))) <= 1e-8
    const = _baseline.crossfit_whittaker_baseline(np.full(n, 3.25), lam)
    assert np.max(np.abs(const - 3.25)) <= 1e-8                                  # D'D annihilates constants
    ramp = _baseline.crossfit_whittaker_baseline(np.arange(n, dtype=np.float64) * 1e-6, lam)
    assert np.max(np.abs(ramp - np.arange(n) * 1e-6)) <= 1e-7                    # ... and straight lines


def test_scores_invariant_to_sample_order_and_sharding_full_chr21():
    import torch
    from rocco_b200 import pipeline
    from rocco_b200.synth import chrom_matrix_torch
    dev = torch.device("cuda", 0)
    x = chrom_matrix_torch(20, 934_200, 21, dev, torch.float64)
    prm = pipeline.score_params(prior_df=6.0)
    a = pipeline.score_loci_wls_device(x, params=prm).clone()
    perm = torch.randperm(20, device=dev)
    b = pipeline.score_loci_wls_device(x[perm].contiguous(), params=prm)
    assert float((a - b).abs().max()) <= 1e-11 * max(1.0, float(a.abs().max()))
    acc = pipeline.score_partial_device(x[:7].contiguous(), prm) + pipeline.score_partial_device(x[7:].contiguous(), prm)
    c = pipeline.score_finalize_device(acc, 20, prm)
    assert float((a - c).abs().max()) <= 1e-11 * max(1.0, float(a.abs().max()))
    # float32 storage of the SAME values gives the same scores (the reference widens to float64 first)
    x32 = x.to(torch.float32)
    d64 = pipeline.score_loci_wls_device(x32.to(torch.float64), params=prm)
    d32 = pipeline.score_loci_wls_device(x32, params=prm)
    assert torch.equal(d64, d32)


def test_sampled_pilot_offset_cancels_at_chr1_size():
    """DESIGN.md section 4, deviation a2: above 4096 bins the row-median pilot offset (inference.py:333) is the median
    of a 4096-point sample.  The offset only enters through  y - B(y)  and the smoother B reproduces constants, so it
    cancels up to rounding.  Bounded here at the longest chromosome (4,979,129 bins) against the EXACT np.median mode."""
    import torch
    from rocco_b200 import pipeline
    from rocco_b200.synth import chrom_bins, chrom_matrix_torch, chrom_seed
    n = chrom_bins("chr1")
    assert n == 4_979_129
    x = chrom_matrix_torch(4, n, chrom_seed("chr1"), torch.device("cuda", 0), torch.float64)
    s0, d0 = pipeline.score_loci_wls_device(x, params=pipeline.score_params(prior_df=6.0), details=True)
    s1, d1 = pipeline.score_loci_wls_device(x, params=pipeline.score_params(prior_df=6.0, exact_pilot=True), details=True)
    # the exact mode really is np.median of log2(max(x,0)+1) per row
    y = torch.log2(torch.clamp(x, min=0.0) + 1.0)
    med = torch.sort(y, dim=1).values[:, [n // 2]]               # n is odd
    c_ref_offset = (d1["centered_matrix"] - (y - med)).abs().max()
    assert float(c_ref_offset) < 10.0                            # (sanity: same scale; the baseline itself is O(1))
    assert float((d0["centered_matrix"] - d1["centered_matrix"]).abs().max()) <= 1e-9
    assert float(((s0 - s1).abs() / torch.clamp(s1.abs(), min=1e-3)).max()) <= 1e-7


def test_exact_pilot_mode_matches_numpy_median():
    import torch
    from oracle import oracle as orc
    from rocco_b200 import pipeline
    from rocco_b200.synth import chrom_matrix_numpy
    x = chrom_matrix_numpy(3, 50_000, seed=9)                    # even length: mean of the two middle values
    _, d = pipeline.score_loci_wls_device(torch.from_numpy(x).cuda(), params=pipeline.score_params(prior_df=6.0, exact_pilot=True), details=True)
    _, want = orc.score_loci_wls(x, prior_df=6.0, return_details=True)
    assert float(np.max(np.abs(d["centered_matrix"].cpu().numpy() - want["centered_matrix"]))) <= 1e-9

"""Generate tests/golden/reference_budget_v1_11_0.npz by running the REAL reference (ROCCO v1.11.0) here.

    python tests/golden/make_golden_budget.py          (needs /root/reference; see make_golden.py for the import recipe)

Covers SURVEY.md 8(f) ranks 1-2: the dependent-wild-bootstrap budget null (inference.py:446-1148) and the automatic
gamma (rocco.py:751-789).  Only small input/output vectors are stored.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import OUT, import_reference  # noqa: E402


def main():
    from rocco_b200.synth import chrom_matrix_numpy
    rocco = import_reference()
    inf, rr = rocco.inference, rocco.rocco
    g = {}
    cases = {
        # the reference's own test input (tests/test_rocco.py:462-500)
        "ref_test": dict(kw=dict(dependence_lag_hint=16, num_null_draws=6)),
        "synth_a": dict(kw=dict(dependence_lag_hint=101, num_null_draws=5, prior_df=6.0, random_seed=3)),
        "synth_b": dict(kw=dict(num_null_draws=25, random_seed=11)),          # default bandwidth rule, adaptive stop allowed
    }
    x = np.arange(512, dtype=np.float64)
    p1 = 6.0 * np.exp(-0.5 * ((x - 120.0) / 15.0) ** 2)
    p2 = 5.5 * np.exp(-0.5 * ((x - 320.0) / 15.0) ** 2)
    mats = {
        "ref_test": np.vstack([0.25 + p1 + p2 + 0.05 * np.sin(x / 13.0), 0.20 + 0.95 * p1 + 1.05 * p2 + 0.04 * np.cos(x / 15.0),
                               0.22 + 1.1 * p1 + 0.9 * p2 + 0.05 * np.sin(x / 17.0)]),
        "synth_a": chrom_matrix_numpy(6, 3000, seed=21),
        "synth_b": chrom_matrix_numpy(4, 6000, seed=5),
    }
    for tag, case in cases.items():
        kw = dict(case["kw"])
        score_kw = {k: kw[k] for k in ("prior_df",) if k in kw}
        scores, det = inf.score_loci_wls(mats[tag], return_details=True, **score_kw)
        frac, meta = inf.estimate_budget_nonnull_fraction_from_wild_bootstrap_null(
            det["centered_matrix"], observed_scores=scores, return_details=True, **kw)
        g[f"{tag}_matrix"] = mats[tag]
        g[f"{tag}_centered"] = det["centered_matrix"]
        g[f"{tag}_scores"] = scores
        g[f"{tag}_kwargs"] = np.array(json.dumps(kw))
        g[f"{tag}_fraction"] = np.array(frac)
        g[f"{tag}_meta"] = np.array(json.dumps({k: (v if isinstance(v, (str, bool)) else float(v)) for k, v in meta.items()}))
        gam, gmeta = rr._resolve_chrom_gamma(tag, {"gamma": None}, scores, meta)
        g[f"{tag}_gamma"] = np.array(gam)
        g[f"{tag}_gamma_meta"] = np.array(json.dumps({k: (v if isinstance(v, str) else float(v)) for k, v in gmeta.items()}))
    # building blocks
    g["kernel_b8"] = inf._build_budget_bootstrap_kernel(8)
    g["kernel_b101"] = inf._build_budget_bootstrap_kernel(101)
    rng = np.random.default_rng(5)
    g["weights_n4000_b16"] = inf._generate_dependent_wild_weights(4000, inf._build_budget_bootstrap_kernel(16), np.random.default_rng(77))
    v = np.convolve(rng.standard_normal(5000), np.ones(9) / 9.0, mode="same") + 0.1 * rng.standard_normal(5000)
    g["ess_values"] = v
    g["ess_out"] = np.array(inf._estimate_effective_sample_size(v, 404), dtype=np.float64)
    g["ess_out_short"] = np.array(inf._estimate_effective_sample_size(v[:40], 64), dtype=np.float64)
    g["bandwidth_rules"] = np.array([[n, -1 if h is None else h, inf._resolve_budget_bootstrap_bandwidth(n, h),
                                      inf._resolve_budget_ess_max_lag(n, h)]
                                     for n in (1, 2, 9, 512, 3000, 934200, 4980000) for h in (None, 16, 25, 101, 1000)], dtype=np.int64)
    # narrowPeak summit offsets (rocco.py:809-872): awkward tracks (NaN / inf means, ties, peaks off the track, empty peaks)
    import tempfile
    rng = np.random.default_rng(9)
    tmp = tempfile.mkdtemp(prefix="summit_")
    chrom_cache, lines = {}, []
    for chrom, nb in (("chrA", 4000), ("chrB", 777)):
        intervals = np.arange(0, 50 * (nb + 1), 50, dtype=np.int64) + 1000
        mean = rng.normal(size=nb + 1)
        mean[rng.integers(0, nb, 60)] = np.nan
        mean[rng.integers(0, nb, 10)] = np.inf
        mean[rng.integers(0, nb, 10)] = -np.inf
        mean[100:140] = 2.5                                   # a tie plateau: first occurrence must win
        mean[300:320] = np.nan                                # a peak with no finite value
        chrom_cache[chrom] = {"summit_track_file": rr._cpy_narrowpeak_summit_track(chrom, intervals, mean)}
        g[f"summit_{chrom}_intervals"], g[f"summit_{chrom}_mean"] = intervals, mean
        starts = np.sort(rng.integers(0, 50 * nb + 3000, 300))
        for s_ in starts:
            lines.append(f"{chrom}\t{int(s_)}\t{int(s_ + rng.integers(0, 900))}")
        lines += [f"{chrom}\t{1000 + 50 * 100}\t{1000 + 50 * 140}", f"{chrom}\t{1000 + 50 * 300}\t{1000 + 50 * 320}",
                  f"{chrom}\t{1000 + 50 * 100 + 7}\t{1000 + 50 * 100 + 8}"]
    lines.append("chrNone\t10\t500")                          # chromosome without a track
    peak_file = os.path.join(tmp, "peaks.bed")
    open(peak_file, "w").write("\n".join(lines) + "\n")
    out = rr._write_narrowpeak_summit_offsets(peak_file, chrom_cache, os.path.join(tmp, "offsets.tsv"))
    g["summit_peaks_text"] = np.array(open(peak_file).read())
    g["summit_offsets_text"] = np.array(open(out).read())
    np.savez_compressed(os.path.join(OUT, "reference_budget_v1_11_0.npz"), **g)
    print("wrote", len(g), "arrays")


if __name__ == "__main__":
    main()

"""Generate tests/golden/reference_eb_budgets_v1_11_0.json by running the REAL reference (ROCCO v1.11.0) here.

    python tests/golden/make_golden_eb.py          (needs /root/reference; see make_golden.py for the import recipe)

Covers the remainder of SURVEY.md 8(f) rank 2: `estimate_empirical_bayes_budgets` / `fit_beta_prior_mle`
(inference.py:1488-1737) and `_resolve_budgets` (rocco.py:1113-1143) -- the step between the per-chromosome budget
estimates and the per-chromosome solves on the reference's default CLI path.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import import_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_eb_budgets_v1_11_0.json")


def main():
    rocco = import_reference()
    inf, rr = rocco.inference, rocco.rocco
    from rocco_b200.synth import HG38_SIZES, HG_PARAMS, chrom_bins
    rng = np.random.default_rng(2024)
    names = list(HG38_SIZES)
    totals24 = {c: float(chrom_bins(c)) for c in names}
    cases = {}
    # (a) 24 chromosomes, rates dispersed around the hg_params budgets -> beta-binomial MLE
    cand = {c: float(np.round(totals24[c] * HG_PARAMS[c][0] * rng.uniform(0.6, 1.5))) for c in names}
    cases["hg38_dispersed"] = dict(cand=cand, tot=totals24, kw={})
    # (b) the same rate everywhere: observed variance at the binomial floor -> boundary prior
    cases["hg38_at_floor"] = dict(cand={c: float(np.round(totals24[c] * 0.03)) for c in names}, tot=totals24, kw={})
    # (c) three chromosomes -> weak pooled prior; (d) one chromosome -> default prior
    three = ["chr19", "chr21", "chrX"]
    cases["three"] = dict(cand={c: cand[c] for c in three}, tot={c: totals24[c] for c in three}, kw={})
    cases["one"] = dict(cand={"chr21": cand["chr21"]}, tot={"chr21": totals24["chr21"]}, kw={})
    # (e) non-default knobs and small counts (posterior quantile matters)
    small = {f"c{k}": float(t) for k, t in enumerate([400, 900, 1500, 2500, 5200, 800])}
    cases["small_counts"] = dict(cand={k: float(np.round(v * r)) for (k, v), r in zip(small.items(), [0.01, 0.08, 0.03, 0.2, 0.05, 0.0])},
                                 tot=small, kw=dict(posterior_quantile=0.25, min_budget=0.002, max_budget=0.15, init_center=0.1, init_strength=4.0))
    out = {}
    for tag, c in cases.items():
        budgets, meta = inf.estimate_empirical_bayes_budgets(c["cand"], c["tot"], **c["kw"])
        out[tag] = {"candidate_counts": c["cand"], "total_counts": c["tot"], "kwargs": c["kw"], "budgets": budgets,
                    "meta": {k: (v if isinstance(v, (str, bool)) else float(v)) for k, v in meta.items()}}
        if tag in ("hg38_dispersed", "small_counts"):
            a, b = inf.fit_beta_prior_mle(np.array(list(c["cand"].values())), np.array(list(c["tot"].values())),
                                          **{k: v for k, v in c["kw"].items() if k in ("init_center", "init_strength")})
            out[tag]["mle"] = [float(a), float(b)]
    # _resolve_budgets over a chromosome cache (rocco.py:1113-1143)
    cache = {c: {"budget_count_hat": cand[c], "total_count": totals24[c]} for c in names}
    res = {}
    for tag, args in {"auto": dict(budget=None, scale_chrom_budgets=1.0, budget_posterior_quantile=0.01),
                      "target_0.03_scaled": dict(budget=0.03, scale_chrom_budgets=1.25, budget_posterior_quantile=0.05)}.items():
        b, meta = rr._resolve_budgets(cache, args)
        res[tag] = {"args": args, "budgets": b, "meta": {k: (v if isinstance(v, (str, bool)) else float(v)) for k, v in meta.items()}}
    out["resolve_budgets"] = {"cache": cache, "cases": res}
    with open(OUT, "w") as fh:
        json.dump(out, fh, indent=1)          # key order IS data: the optimiser's path depends on the summation order
    print("wrote", OUT)


if __name__ == "__main__":
    main()

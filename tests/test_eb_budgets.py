"""Empirical-Bayes chromosome budgets and `_resolve_budgets` against golden values produced by the REAL reference
(tests/golden/make_golden_eb.py -> reference_eb_budgets_v1_11_0.json): SURVEY.md section 8(f) rank 2.  Host SciPy on
both sides, so these run without a GPU."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def eb():
    return json.load(open(os.path.join(HERE, "golden", "reference_eb_budgets_v1_11_0.json")))


def _close(a, b, tol=1e-9):
    return abs(a - b) <= tol * max(1.0, abs(b))


@pytest.mark.parametrize("tag", ["hg38_dispersed", "hg38_at_floor", "three", "one", "small_counts"])
def test_empirical_bayes_budgets_match_reference(eb, tag):
    from rocco_b200.inference import estimate_empirical_bayes_budgets, fit_beta_prior_mle
    case = eb[tag]
    budgets, meta = estimate_empirical_bayes_budgets(case["candidate_counts"], case["total_counts"], **case["kwargs"])
    assert list(budgets) == list(case["candidate_counts"])
    assert meta["prior_fit_method"] == case["meta"]["prior_fit_method"]
    assert meta["posterior_summary"] == "beta_quantile"
    for k, v in case["meta"].items():
        if isinstance(v, (str, bool)):
            assert meta[k] == v, k
        else:
            assert _close(meta[k], v), (k, meta[k], v)
    for c, b in case["budgets"].items():
        assert _close(budgets[c], b), (c, budgets[c], b)
    if "mle" in case:
        a, b = fit_beta_prior_mle(np.array(list(case["candidate_counts"].values())), np.array(list(case["total_counts"].values())),
                                  **{k: v for k, v in case["kwargs"].items() if k in ("init_center", "init_strength")})
        assert _close(a, case["mle"][0]) and _close(b, case["mle"][1])


def test_empirical_bayes_budgets_argument_checks():
    from rocco_b200.inference import estimate_empirical_bayes_budgets, fit_beta_prior_mle
    with pytest.raises(ValueError):
        estimate_empirical_bayes_budgets({"a": 1.0, "b": 2.0}, {"b": 10.0, "a": 10.0})
    with pytest.raises(ValueError):
        estimate_empirical_bayes_budgets({"a": 1.0}, {"a": 10.0}, posterior_quantile=1.0)
    with pytest.raises(ValueError):
        fit_beta_prior_mle(np.zeros(3), np.zeros(2))
    assert fit_beta_prior_mle(np.zeros(0), np.zeros(0)) == (1.0, 1.0)


@pytest.mark.parametrize("tag", ["auto", "target_0.03_scaled"])
def test_resolve_budgets_matches_reference(eb, tag):
    from rocco_b200.rocco import _resolve_budgets
    case = eb["resolve_budgets"]["cases"][tag]
    budgets, meta = _resolve_budgets(eb["resolve_budgets"]["cache"], case["args"])
    assert list(budgets) == list(case["budgets"])
    for c, b in case["budgets"].items():
        assert 0.005 <= budgets[c] <= 0.1
        assert _close(budgets[c], b), (c, budgets[c], b)
    assert _close(meta["alpha"], case["meta"]["alpha"]) and _close(meta["beta"], case["meta"]["beta"])

"""GPU parity: column-wise statistics over the sample axis (rocco.py:243-355) vs golden vectors / the oracle."""
import numpy as np
import pytest

from rocco_b200.synth import chrom_matrix_numpy

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rb():
    import rocco_b200
    return rocco_b200


def test_column_statistics_match_reference_golden(rb, golden):
    x = golden["col_x"]
    assert np.array_equal(rb.score_central_tendency_chrom(x), golden["col_median"])
    assert np.array_equal(rb.score_central_tendency_chrom(x, method="quantile", quantile=0.75), golden["col_q75"])
    assert np.allclose(rb.score_central_tendency_chrom(x, method="quantile", quantile=0.25, power=0.5), golden["col_q25_pow"], rtol=1e-14, atol=0)
    assert np.allclose(rb.score_central_tendency_chrom(x, method="tmean", tprop=0.1), golden["col_tmean"], rtol=1e-14, atol=0)
    assert np.array_equal(rb.score_central_tendency_chrom(x, method="mean"), golden["col_mean"])
    assert np.array_equal(rb.score_dispersion_chrom(x, method="mad"), golden["col_mad"])
    assert np.allclose(rb.score_dispersion_chrom(x, method="iqr"), golden["col_iqr"], rtol=1e-14, atol=1e-300)
    assert np.allclose(rb.score_dispersion_chrom(x, method="std"), golden["col_std"], rtol=1e-14, atol=0)
    x10 = golden["col10_x"]
    assert np.array_equal(rb.score_central_tendency_chrom(x10), golden["col10_median"])
    assert np.array_equal(rb.score_central_tendency_chrom(x10, method="quantile", quantile=0.75), golden["col10_q75"])
    assert np.array_equal(rb.score_dispersion_chrom(x10, method="mad"), golden["col10_mad"])
    assert np.allclose(rb.score_dispersion_chrom(x10, method="iqr", rng=(10, 90)), golden["col10_iqr"], rtol=1e-14, atol=1e-300)
    # reference tests/test_rocco.py:895-896 (bigWig path: median of two samples)
    bw = rb.score_central_tendency_chrom(np.array([[0.0, 2.0, 1.0, 0.0], [0.0, 3.0, 2.0, 0.0]]))
    assert bw.tolist() == [0.0, 2.5, 1.5, 0.0]


@pytest.mark.parametrize("m,n,seed", [(2, 1000, 0), (3, 777, 1), (10, 50_000, 2), (11, 50_001, 3), (100, 20_000, 4), (129, 3_000, 5), (300, 2_000, 6)])
def test_column_statistics_match_oracle(rb, oracle, m, n, seed):
    x = chrom_matrix_numpy(m, n, seed=seed)
    for kw in (dict(), dict(method="quantile", quantile=0.75), dict(method="quantile", quantile=0.25),
               dict(method="quantile", quantile=0.5, power=2.0), dict(method="mean"), dict(method="t-mean", tprop=0.05),
               dict(method="tmean", tprop=0.2)):
        want = oracle.score_central_tendency_chrom(x, **kw)
        got = rb.score_central_tendency_chrom(x, **kw)
        if kw.get("method", "quantile") == "quantile" and kw.get("power", 1.0) == 1.0 or kw.get("method") == "mean":
            assert np.array_equal(got, want), kw                  # order statistics / sequential sums: exact
        else:
            assert np.allclose(got, want, rtol=1e-13, atol=0), kw
    for kw in (dict(method="mad"), dict(method="iqr"), dict(method="IQR", rng=(10, 90)), dict(method="std"),
               dict(method="tstd", tprop=0.1)):
        want = oracle.score_dispersion_chrom(x, **kw)
        got = rb.score_dispersion_chrom(x, **kw)
        if kw["method"] == "mad":
            assert np.array_equal(got, want), kw
        else:
            assert np.allclose(got, want, rtol=1e-12, atol=1e-300, equal_nan=True), kw


def test_single_sample_and_errors(rb):
    x = np.array([[1.0, 4.0, 9.0]])
    assert rb.score_central_tendency_chrom(x, power=0.5).tolist() == [1.0, 2.0, 3.0]
    assert rb.score_dispersion_chrom(x).tolist() == [0.0, 0.0, 0.0]
    with pytest.raises(ValueError):
        rb.score_central_tendency_chrom(np.zeros((2, 3)), method="nope")
    with pytest.raises(ValueError):
        rb.score_dispersion_chrom(np.zeros((2, 3)), method="nope")
    with pytest.raises(ValueError):
        rb.score_central_tendency_chrom(np.zeros(3))

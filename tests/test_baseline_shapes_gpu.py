"""Parity against the oracle (the reference's own kernels where oracle/_ref exists) AT the shapes BASELINE.json names.

  config 1   chr21: 934,200 bins x 10 and x 100 samples, budget 0.02, gamma 1.0 (hg_params.csv)   SURVEY.md section 8d
  config 2   chr19 (1,172,353 bins) + chrX (3,120,818) x 10 samples through combine_chrom_results
  config 3   a three-chromosome slice of the genome-wide run (chr20, chr21, chr22 x 10) through the device pipeline

Gates: masks / counts / intervals / BED text identical; scores within 1e-5 relative (denominator
max(|ref|, 1e-3)); objective within 1e-6 relative; the searched multiplier is reported against the oracle's.
Both sides see the same bytes: the matrix comes from the seeded NumPy generator of rocco_b200.synth.
"""
import os

import numpy as np
import pytest

from tests.conftest import rel_err
from rocco_b200.synth import HG_PARAMS, chrom_bins, chrom_matrix_numpy, chrom_seed

pytestmark = pytest.mark.gpu

TOL = 1e-5
STEP = 50
PRIOR_DF = 6.0


@pytest.fixture(scope="module")
def rb():
    import rocco_b200
    return rocco_b200


def _kind(oracle):
    return "reference" if oracle.reference_available() else "port"


def _oracle_chrom(oracle, name, x, workdir):
    """the reference's path for one chromosome: scores -> budgeted solve -> BED file"""
    kind = _kind(oracle)
    budget, gamma = HG_PARAMS[name]
    scores = oracle.score_loci_wls(x, prior_df=PRIOR_DF, kind=kind)
    sol, obj, det = oracle.solve_chrom_exact(scores, budget=budget, gamma=gamma, return_details=True, kind=kind)
    cwd = os.getcwd()
    os.chdir(workdir)
    try:
        bed = oracle.chrom_solution_to_bed(name, np.arange(0, STEP * x.shape[1], STEP), sol, ID="want")
        bed = os.path.abspath(bed)
    finally:
        os.chdir(cwd)
    return scores, sol, obj, det, bed


def _ours_chrom(rb, name, x, workdir):
    budget, gamma = HG_PARAMS[name]
    scores = rb.score_loci_wls(x, prior_df=PRIOR_DF)
    sol, obj, det = rb.solve_chrom_exact(scores, budget=budget, gamma=gamma, return_details=True)
    cwd = os.getcwd()
    os.chdir(workdir)
    try:
        bed = os.path.abspath(rb.chrom_solution_to_bed(name, np.arange(0, STEP * x.shape[1], STEP), sol, ID="got"))
    finally:
        os.chdir(cwd)
    return scores, sol, obj, det, bed


def _check_chrom(got, want, name):
    gs, gsol, gobj, gdet, gbed = got
    ws, wsol, wobj, wdet, wbed = want
    e = rel_err(gs, ws)
    assert e <= TOL, (name, e)
    assert np.array_equal(gsol, wsol), (name, int(np.sum(gsol != wsol)))
    assert gdet["selected_count"] == wdet["selected_count"]
    assert abs(gobj - wobj) <= 1e-6 * abs(wobj), (name, gobj, wobj)
    # the multiplier is searched on scores that differ from the reference's by ~1e-7 relative, so it agrees to that level
    assert abs(gdet["selection_penalty"] - wdet["selection_penalty"]) <= 1e-5 * max(1.0, abs(wdet["selection_penalty"]))
    assert open(gbed, "rb").read() == open(wbed, "rb").read(), name


@pytest.mark.parametrize("m", [10, 100])
def test_config1_chr21_full_size(rb, oracle, tmp_path, m):
    n = chrom_bins("chr21", STEP)
    assert n == 934_200
    x = chrom_matrix_numpy(m, n, seed=chrom_seed("chr21"))
    want = _oracle_chrom(oracle, "chr21", x, str(tmp_path))
    got = _ours_chrom(rb, "chr21", x, str(tmp_path))
    _check_chrom(got, want, f"chr21 x {m}")
    assert want[3]["selected_count"] <= int(np.floor(n * 0.02))


def test_config1_the_same_multiplier_on_the_same_scores(rb, oracle):
    """search parity proper: both implementations search the SAME score vector (the oracle's) at chr21 size"""
    n = chrom_bins("chr21", STEP)
    x = chrom_matrix_numpy(10, n, seed=chrom_seed("chr21"))
    scores = oracle.score_loci_wls(x, prior_df=PRIOR_DF, kind=_kind(oracle))
    wsol, wobj, wdet = oracle.solve_chrom_exact(scores, budget=0.02, gamma=1.0, return_details=True, kind=_kind(oracle))
    gsol, gobj, gdet = rb.solve_chrom_exact(scores, budget=0.02, gamma=1.0, return_details=True)
    assert np.array_equal(gsol, wsol)
    assert gdet["selected_count"] == wdet["selected_count"]
    assert abs(gdet["selection_penalty"] - wdet["selection_penalty"]) <= 1e-9
    assert abs(gobj - wobj) <= 1e-6 * abs(wobj)


def test_config2_chr19_chrX_combined(rb, oracle, tmp_path):
    want_files, got_files = [], []
    (tmp_path / "w").mkdir()
    (tmp_path / "g").mkdir()
    for name in ("chr19", "chrX"):
        n = chrom_bins(name, STEP)
        assert n == {"chr19": 1_172_353, "chrX": 3_120_818}[name]
        x = chrom_matrix_numpy(10, n, seed=chrom_seed(name))
        want = _oracle_chrom(oracle, name, x, str(tmp_path / "w"))
        got = _ours_chrom(rb, name, x, str(tmp_path / "g"))
        _check_chrom(got, want, name)
        want_files.append(want[4])
        got_files.append(got[4])
    w = oracle.combine_chrom_results(want_files, str(tmp_path / "want_combined.bed"))
    g = rb.combine_chrom_results(got_files, str(tmp_path / "got_combined.bed"))
    assert open(g, "rb").read() == open(w, "rb").read()


def test_config3_slice_through_the_device_pipeline(rb, oracle, tmp_path):
    """chr20 + chr21 + chr22 x 10 with hg_params budgets/gammas: the batched device path (one launch set per stage over
    the shard, the path bench.py times) against the oracle chromosome by chromosome"""
    import torch
    from rocco_b200 import pipeline
    names = ["chr20", "chr21", "chr22"]
    mats = [chrom_matrix_numpy(10, chrom_bins(c, STEP), seed=chrom_seed(c)) for c in names]
    budgets = [HG_PARAMS[c][0] for c in names]
    gammas = [HG_PARAMS[c][1] for c in names]
    d_mats = [torch.from_numpy(x).cuda() for x in mats]
    shard = pipeline.run_shard(d_mats, budgets, gammas, params=pipeline.score_params(prior_df=PRIOR_DF))
    got_text = pipeline.runs_to_bed_text(names, shard["runs"], STEP)
    want_files = []
    for c, x, res in zip(names, mats, shard["results"]):
        ws, wsol, wobj, wdet, wbed = _oracle_chrom(oracle, c, x, str(tmp_path))
        assert res["selected_count"] == wdet["selected_count"], c
        assert abs(res["objective"] - wobj) <= 1e-6 * abs(wobj), c
        want_files.append(wbed)
    want = oracle.combine_chrom_results(want_files, str(tmp_path / "want.bed"))
    assert got_text.encode() == open(want, "rb").read()

"""Generate tests/golden/*.npz by running the REAL reference (ROCCO v1.11.0) in the build container.

Run once where /root/reference is mounted:   python tests/golden/make_golden.py
The reference package is assembled in a temp dir out of symlinks to its own .py files plus the
extension binaries compiled by `make -C oracle ref`, with a 2-class stub for the (absent, unused
on this path) `pysam` import.  Nothing of the reference is copied into this repo; only the small
input/output vectors below are committed, so the GPU box (which has no /root/reference) can check
the oracle and the CUDA path against them.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/rocco"
OUT = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, REPO)


def import_reference():
    import subprocess
    subprocess.check_call(["make", "-s", "-C", os.path.join(REPO, "oracle"), "ref"])
    tmp = tempfile.mkdtemp(prefix="rocco_ref_")
    pkg = os.path.join(tmp, "rocco")
    os.makedirs(pkg)
    for f in os.listdir(REF):
        if f.endswith((".py", ".sizes", ".csv")):
            os.symlink(os.path.join(REF, f), os.path.join(pkg, f))
    refdir = os.path.join(REPO, "oracle", "_ref")
    for f in os.listdir(refdir):
        os.symlink(os.path.join(refdir, f), os.path.join(pkg, f))
    stubs = os.path.join(tmp, "stubs")
    os.makedirs(stubs)
    with open(os.path.join(stubs, "pysam.py"), "w") as fh:
        fh.write("class AlignedSegment: pass\nclass AlignmentFile: pass\n")
    sys.path[:0] = [tmp, stubs]
    import rocco  # noqa: F401
    import rocco.dp
    import rocco.inference
    import rocco.rocco
    assert rocco.__version__ == "1.11.0", rocco.__version__
    return rocco


def main():
    from rocco_b200.synth import chrom_matrix_numpy

    rocco = import_reference()
    inf, dp, rr = rocco.inference, rocco.dp, rocco.rocco
    g: dict[str, np.ndarray] = {}

    # --- scoring: synthetic 6 x 3000 (generator of SURVEY.md 8d), default and CLI-style parameters
    x = chrom_matrix_numpy(6, 3000, seed=21)
    g["score_x"] = x
    for tag, kw in (("a", {}), ("b", dict(prior_df=6.0, lower_bound_z=0.5, precision_floor_ratio=0.05)),
                    ("c", dict(min_effect=0.25))):
        sc, det = inf.score_loci_wls(x, return_details=True, **kw)
        g[f"score_{tag}_scores"] = sc
        for k in ("mean", "raw_variance", "prior_variance", "moderated_variance", "standard_error", "z_scores"):
            g[f"score_{tag}_{k}"] = det[k]
        if tag == "a":
            g["score_a_centered"] = det["centered_matrix"]
            g["score_a_meta"] = np.array([det["local_baseline_window"], det["local_baseline_lambda"],
                                          det["prior_spatial_window"], det["degrees_of_freedom"][0]])
    # float32 input, as --low_memory feeds it (readtracks.py:621)
    x32 = x.astype(np.float32)
    g["score_f32_scores"] = inf.score_loci_wls(x32)

    # --- small-n branches (n < 25: zero baseline; n < 5: robust-scale variances)
    for tag, arr in (("n2", np.array([[1.0, 15.0]])),
                     ("n3", np.array([[1.0, 3.0, 7.0], [1.2, 2.8, 6.5]])),
                     ("n4", np.array([[0.05, 1.0, 1.0, 0.05], [0.04, 1.0, 1.0, 0.04], [0.06, 1.0, 1.0, 0.06]])),
                     ("n24", chrom_matrix_numpy(3, 24, seed=3)),
                     ("n25", chrom_matrix_numpy(3, 25, seed=4)),
                     ("n130", chrom_matrix_numpy(2, 130, seed=5))):
        sc, det = inf.score_loci_wls(arr, lower_bound_z=0.0, return_details=True)
        g[f"small_{tag}_x"] = arr
        g[f"small_{tag}_scores"] = sc
        g[f"small_{tag}_mean"] = det["mean"]
        g[f"small_{tag}_se"] = det["standard_error"]

    # --- stage-level: baseline and centered WLS on their own
    rng = np.random.default_rng(11)
    y = rng.normal(size=(3, 700)).cumsum(axis=1) * 0.05 + rng.normal(size=(3, 700))
    g["base_y"] = y
    g["base_out"] = inf._estimate_local_background_matrix(y)[0]
    t = np.arange(129, dtype=np.float64)
    y129 = 2.5 * np.exp(-0.5 * ((t - 64.0) / 18.0) ** 2) + 5.0 * np.exp(-0.5 * ((t - 64.0) / 2.5) ** 2)
    g["base129_y"] = y129
    g["base129_out"] = inf._consenrich_crossfit_whittaker_baseline(y129, block_size=41)
    c = rng.normal(size=(4, 1500)) * (0.3 + np.abs(np.sin(np.arange(1500) / 90.0)))
    g["wls_centered"] = c
    sc, det = inf._score_centered_wls_matrix(c, prior_df=6.0, spatial_window=31)
    g["wls_scores"] = sc
    for k in ("mean", "raw_variance", "prior_variance", "moderated_variance", "standard_error"):
        g[f"wls_{k}"] = det[k]

    # --- chain DP / multiplier search
    scores = g["score_a_scores"]
    for tag, budget, gamma in (("g1", 0.02, 1.0), ("g7", 0.03, 6.86), ("g0", 0.1, 0.0)):
        sol, obj, det = dp.solve_chrom_exact(scores, budget=budget, gamma=gamma, return_details=True)
        g[f"dp_{tag}_mask"] = sol
        g[f"dp_{tag}_meta"] = np.array([budget, gamma, obj, det["penalized_objective"],
                                        det["selected_count"], det["selection_penalty"]])
    r7 = np.random.default_rng(7)
    bs, bc = r7.normal(size=9), r7.uniform(0.2, 1.3, size=8)
    g["dp_bf_scores"], g["dp_bf_costs"] = bs, bc
    for k, pen in enumerate((-0.5, 0.0, 0.6, 1.4)):
        sol, val, cnt = dp.solve_penalized_chain(bs, bc, pen)
        g[f"dp_bf_{k}_mask"] = sol
        g[f"dp_bf_{k}_meta"] = np.array([pen, val, cnt])
    s8 = np.array([0.5, 1.5, 1.4, -0.2, 3.0, 2.8, -0.1, 0.1])
    sol, obj, det = dp.solve_chrom_exact(s8, budget=0.375, gamma=1.0, return_details=True)
    g["dp_s8_scores"], g["dp_s8_mask"] = s8, sol
    g["dp_s8_meta"] = np.array([obj, det["penalized_objective"], det["selected_count"], det["selection_penalty"]])
    # ties: constant and integer-valued scores exercise the (value, fewer-count) tie-break
    for tag, arr in (("const", np.full(40, -1.0)), ("ints", np.array([1, -1, 1, -1, 2, -2, 0, 0, 1, 1, -1, 3.0] * 5))):
        for k, (b, gm) in enumerate(((0.25, 1.0), (0.5, 0.5), (None, 1.0))):
            sol, obj, det = dp.solve_chrom_exact(arr, budget=b, gamma=gm, return_details=True)
            g[f"dp_tie_{tag}_{k}_mask"] = sol
            g[f"dp_tie_{tag}_{k}_meta"] = np.array([-1.0 if b is None else b, gm, obj, det["penalized_objective"],
                                                    det["selected_count"], det["selection_penalty"]])
        g[f"dp_tie_{tag}_scores"] = arr

    # --- column statistics (rocco.py:243-355)
    cx = chrom_matrix_numpy(11, 400, seed=9)
    g["col_x"] = cx
    g["col_median"] = rr.score_central_tendency_chrom(cx)
    g["col_q75"] = rr.score_central_tendency_chrom(cx, method="quantile", quantile=0.75)
    g["col_q25_pow"] = rr.score_central_tendency_chrom(cx, method="quantile", quantile=0.25, power=0.5)
    g["col_tmean"] = rr.score_central_tendency_chrom(cx, method="tmean", tprop=0.1)
    g["col_mean"] = rr.score_central_tendency_chrom(cx, method="mean")
    g["col_mad"] = rr.score_dispersion_chrom(cx, method="mad")
    g["col_iqr"] = rr.score_dispersion_chrom(cx, method="iqr")
    g["col_std"] = rr.score_dispersion_chrom(cx, method="std")
    cx10 = chrom_matrix_numpy(10, 300, seed=10)
    g["col10_x"] = cx10
    g["col10_median"] = rr.score_central_tendency_chrom(cx10)
    g["col10_q75"] = rr.score_central_tendency_chrom(cx10, method="quantile", quantile=0.75)
    g["col10_mad"] = rr.score_dispersion_chrom(cx10, method="mad")
    g["col10_iqr"] = rr.score_dispersion_chrom(cx10, method="iqr", rng=(10, 90))
    g["col_bw"] = rr.score_central_tendency_chrom(np.array([[0.0, 2.0, 1.0, 0.0], [0.0, 3.0, 2.0, 0.0]]))

    # --- mask -> BED text (rocco.py:139-191): last bin dropped, merge, min length
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td:
        os.chdir(td)
        try:
            iv = np.arange(1000, 1000 + 50 * 3000, 50)
            for tag, ml in (("all", None), ("min150", 150)):
                f = rr.chrom_solution_to_bed("chr21", iv, g["dp_g1_mask"], ID="gold", min_length_bp=ml)
                g[f"bed_{tag}_text"] = np.frombuffer(open(f, "rb").read(), dtype=np.uint8)
            toy = np.array([0, 1, 1, 0, 0, 1, 0, 1, 1, 1], dtype=np.uint8)
            f = rr.chrom_solution_to_bed("chrT", np.arange(0, 500, 50), toy)
            g["bed_toy_text"] = np.frombuffer(open(f, "rb").read(), dtype=np.uint8)
        finally:
            os.chdir(cwd)
    g["bed_iv"] = iv

    np.savez_compressed(os.path.join(OUT, "reference_v1_11_0.npz"), **g)
    print("wrote", os.path.join(OUT, "reference_v1_11_0.npz"), "with", len(g), "arrays")


if __name__ == "__main__":
    main()

"""Device-side matrix assembly (rocco_b200.readtracks) against outputs of the REAL reference's
readtracks.py (tests/golden/reference_assembly_v1_11_0.npz) and against the oracle's counting loop."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
CHROM_SIZE = 1_000_003


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "reference_assembly_v1_11_0.npz"))


def case_table():
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from make_golden_assembly import cases
    return cases()


def samples_of(gold, tag, nsamples, paired, kw):
    out = []
    ext = int(kw.get("extend_reads", -1))
    for k in range(nsamples):
        s = {n: gold[f"{tag}_s{k}_{n}"] for n in ("pos", "end", "flag", "mapq", "isize", "mate_same_tid")}
        s.update(read_length=50, resolved_extend_bp=ext if ext > 0 else 0, paired_end_mode=paired,
                 norm_scale=float(gold[f"{tag}_s{k}_norm_scale"]))
        out.append(s)
    return out


@pytest.mark.parametrize("tag", ["single_end", "extended_f32", "centered", "paired_end", "gapped"])
def test_device_assembly_reproduces_reference_matrix(gold, tag):
    from rocco_b200 import readtracks as rt
    step, samples, paired, kw = case_table()[tag]
    intervals, matrix = rt.generate_chrom_matrix_from_reads(
        samples_of(gold, tag, len(samples), paired, kw), CHROM_SIZE, step, const_scale=float(kw.get("const_scale", 1.0)),
        scale_by_step=bool(kw.get("scale_by_step", False)), center_reads=bool(kw.get("center_reads", False)),
        low_memory=bool(kw.get("low_memory", False)))
    want = gold[f"{tag}_matrix"]
    assert np.array_equal(intervals, gold[f"{tag}_intervals"])
    got = matrix.cpu().numpy()
    assert got.dtype == want.dtype and got.shape == want.shape
    assert np.array_equal(got, want), int(np.sum(got != want))          # counts are integers, the scaling is IEEE: bit-exact


def test_device_counts_match_oracle_counter_with_shifts_and_filters(oracle):
    """options the reference's Python never sets (strand shifts, include mask, template-length window) against the
    oracle's restatement of ccounts_backend.c:2416-2574"""
    from oracle import assembly as asm
    from rocco_b200 import readtracks as rt
    for paired, one_per_bin, ext in ((False, 0, 0), (False, 1, 0), (False, 0, 150), (True, 0, 0)):
        r = asm.synthetic_reads(120_000, 2_000_000, seed=31 + one_per_bin + ext, paired=paired)
        start, stop, step = 10_000, 1_999_990, 25
        opt = asm.CountOptions(flag_include=1 if paired else 0, flag_exclude=1796, min_mapping_quality=5, paired_end_mode=int(paired),
                               one_read_per_bin=one_per_bin, read_length=50, min_template_length=120, max_insert_size=450,
                               shift_forward=4, shift_reverse=5, extend_bp=ext)
        want = asm.count_alignment_region(r, start, stop, step, opt)
        got = rt.count_alignment_region(r.pos, r.end, r.flag, r.mapq, r.isize, r.mate_same_tid, start, stop, step, 50,
                                        one_read_per_bin=one_per_bin, flag_include=1 if paired else 0, flag_exclude=1796, extend_bp=ext,
                                        paired_end_mode=int(paired), min_mapping_quality=5, min_template_length=120, max_insert_size=450,
                                        shift_forward_strand53=4, shift_reverse_strand53=5).cpu().numpy()
        assert got.dtype == np.float32 and np.array_equal(got, want)
        assert want.sum() > 0


def test_assembled_matrix_feeds_the_scoring_path(gold, oracle):
    """the device-born matrix goes straight into score_loci_wls on the device: same scores as from the reference's matrix.
    (Columns where every sample has coverage: over the stretches where a sample has no reads at all its centred signal
    is pure rounding noise and the variance trend there is ill-conditioned in ANY implementation.)"""
    from rocco_b200 import pipeline, readtracks as rt
    step, samples, paired, kw = case_table()["single_end"]
    intervals, matrix = rt.generate_chrom_matrix_from_reads(samples_of(gold, "single_end", len(samples), paired, kw), CHROM_SIZE, step)
    cols = np.flatnonzero((intervals >= 300_000) & (intervals < 900_000))
    sub = matrix[:, int(cols[0]):int(cols[-1]) + 1].contiguous()
    got = pipeline.score_loci_wls_device(sub, params=pipeline.score_params(prior_df=6.0)).cpu().numpy()
    want = oracle.score_loci_wls(gold["single_end_matrix"][:, cols[0]:cols[-1] + 1], prior_df=6.0)
    assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-3)) <= 1e-5


def test_assembly_edge_cases():
    import torch
    from rocco_b200 import readtracks as rt
    dev = torch.device("cuda", 0)
    zero = torch.zeros(1000, dtype=torch.float32, device=dev)
    assert rt.assemble_chrom_matrix([{"counts": zero, "count_start": 0, "norm_scale": 1.0}], 50) == (None, None)
    one = zero.clone()
    one[17] = 3.0
    iv, m = rt.assemble_chrom_matrix([{"counts": zero, "count_start": 0, "norm_scale": 1.0},
                                      {"counts": one, "count_start": 500, "norm_scale": 0.5, "const_scale": -1.0}], 50)
    assert iv.tolist() == [500 + 17 * 50] and m.shape == (1, 1) and float(m[0, 0]) == 1.5      # const_scale < 0: not applied
    with pytest.raises(ValueError):
        rt.assemble_chrom_matrix([{"counts": one, "count_start": 7, "norm_scale": 1.0}], 50)

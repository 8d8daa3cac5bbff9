"""Pins the oracle's restatement of the matrix-assembly step (oracle/assembly.py) to outputs of the
REAL reference's readtracks.py (tests/golden/make_golden_assembly.py).  CPU only."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CHROM_SIZE = 1_000_003


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "reference_assembly_v1_11_0.npz"))


def case_table():
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from make_golden_assembly import cases
    return cases()


def oracle_tracks(gold, tag, step, nsamples, paired, kw, asm):
    tracks = []
    for k in range(nsamples):
        r = asm.Reads(*[gold[f"{tag}_s{k}_{n}"] for n in ("pos", "end", "flag", "mapq", "isize", "mate_same_tid")])
        keep = (r.flag & 3844) == 0
        start, end = asm.count_window(int(r.pos[keep].min()), int(r.end[keep].max()), CHROM_SIZE, step)
        ext = int(kw.get("extend_reads", -1))
        opt = asm.CountOptions(flag_include=0, flag_exclude=3844, min_mapping_quality=10, paired_end_mode=int(paired),
                               one_read_per_bin=int(bool(kw.get("center_reads", False))), read_length=50, min_template_length=-1,
                               max_insert_size=1000, shift_forward=0, shift_reverse=0, extend_bp=ext if ext > 0 else 0)
        counts = asm.count_alignment_region(r, start, end, step, opt)
        tracks.append(asm.track_from_counts(counts, start, step, float(gold[f"{tag}_s{k}_norm_scale"]),
                                            scale_by_step=bool(kw.get("scale_by_step", False)),
                                            const_scale=float(kw.get("const_scale", 1.0))))
    return tracks


@pytest.mark.parametrize("tag", ["single_end", "extended_f32", "centered", "paired_end", "gapped"])
def test_oracle_assembly_reproduces_reference(gold, oracle, tag):
    from oracle import assembly as asm
    step, samples, paired, kw = case_table()[tag]
    tracks = oracle_tracks(gold, tag, step, len(samples), paired, kw, asm)
    intervals, matrix = asm.assemble_matrix(tracks, low_memory=bool(kw.get("low_memory", False)))
    assert np.array_equal(intervals, gold[f"{tag}_intervals"])
    assert matrix.dtype == gold[f"{tag}_matrix"].dtype and np.array_equal(matrix, gold[f"{tag}_matrix"])
    if tag == "gapped":
        assert np.any(np.diff(intervals) != step)          # the union really has a hole

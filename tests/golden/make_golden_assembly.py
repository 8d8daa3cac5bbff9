"""Generate tests/golden/reference_assembly_v1_11_0.npz by running the REAL reference's `rocco/readtracks.py` here.

    python tests/golden/make_golden_assembly.py          (needs /root/reference; see make_golden.py for the import recipe)

SURVEY.md 8(f) rank 3 (matrix assembly / input staging).  The reference's BAM counter `_hts_counts` needs htslib, which
cannot be built in this container, so `readtracks._hts_counts` and `_get_bam_count_metadata` are replaced by stand-ins
that serve SYNTHETIC decoded reads through the oracle's restatement of the counting loop (oracle/c/oracle_counts.c).
Everything else is the reference's own code running for real: the count window, scaling, trimming and rounding of
`get_bam_chrom_reads` (readtracks.py:455-518) and the interval union + scatter of `generate_chrom_matrix` (590-633).
Stored: the reads of every sample, the options, and the reference's (intervals, matrix) outputs.
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import import_reference  # noqa: E402

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_assembly_v1_11_0.npz")

CHROM, CHROM_SIZE = "chrT", 1_000_003


def cases():
    return {
        # tag: (step, samples as (n_reads, seed, sub-range of the chromosome the reads fall in), paired, kwargs of generate_chrom_matrix)
        "single_end": (50, [(60_000, 1, (0, CHROM_SIZE)), (40_000, 2, (120_000, 900_000)), (30_000, 3, (300_000, CHROM_SIZE))], False,
                       dict(extend_reads=-1, center_reads=False, low_memory=False)),
        "extended_f32": (20, [(50_000, 4, (0, 500_000)), (50_000, 5, (400_000, CHROM_SIZE)), (200, 6, (10_000, 20_000))], False,
                         dict(extend_reads=180, center_reads=False, low_memory=True, scale_by_step=True)),
        "centered": (50, [(50_000, 7, (5_000, 990_000)), (45_000, 8, (0, CHROM_SIZE))], False,
                     dict(extend_reads=-1, center_reads=True, low_memory=False, const_scale=2.5)),
        "paired_end": (50, [(80_000, 9, (0, CHROM_SIZE)), (60_000, 10, (200_000, 800_000))], True,
                       dict(extend_reads=-1, center_reads=False, low_memory=False)),
        # disjoint read ranges: the union of the interval grids has a gap (columns are only the bins some sample covers)
        "gapped": (50, [(20_000, 11, (0, 200_000)), (20_000, 12, (600_000, 800_000))], False,
                   dict(extend_reads=-1, center_reads=False, low_memory=False)),
    }


def main():
    rocco = import_reference()
    rt = rocco.readtracks
    from oracle import assembly as asm
    store = {}
    reads_by_file = {}
    meta_by_file = {}

    def fake_range(bam_file, chromosome, chrom_size, thread_count=1, flag_exclude=0):
        r = reads_by_file[bam_file]
        keep = (r.flag & flag_exclude) == 0
        return int(r.pos[keep].min()), int(r.end[keep].max())

    def fake_count(bam_file, chromosome, start, end, step, read_length, one_read_per_bin=0, thread_count=1, flag_include=0,
                   flag_exclude=0, extend_bp=0, paired_end_mode=0, min_mapping_quality=0, count_mode="coverage"):
        opt = asm.CountOptions(flag_include=flag_include, flag_exclude=flag_exclude, min_mapping_quality=min_mapping_quality,
                               paired_end_mode=paired_end_mode, one_read_per_bin=one_read_per_bin, read_length=read_length,
                               min_template_length=-1, max_insert_size=1000, shift_forward=0, shift_reverse=0, extend_bp=extend_bp)
        return asm.count_alignment_region(reads_by_file[bam_file], start, end, step, opt)

    rt._hts_counts = types.SimpleNamespace(get_alignment_chrom_range=fake_range, count_alignment_region=fake_count)
    rt._get_bam_count_metadata = lambda bam_file, **kw: meta_by_file[bam_file]
    with tempfile.TemporaryDirectory() as tmp:
        sizes = os.path.join(tmp, "t.sizes")
        open(sizes, "w").write(f"{CHROM}\t{CHROM_SIZE}\n")
        for tag, (step, samples, paired, kw) in cases().items():
            files = []
            for k, (n_reads, seed, (lo, hi)) in enumerate(samples):
                f = os.path.join(tmp, f"{tag}_{k}.bam")
                open(f, "w").write("")
                r = asm.synthetic_reads(n_reads, hi - lo, seed, paired=paired)
                r.pos += lo
                r.end += lo
                reads_by_file[f] = r
                ext = int(kw.get("extend_reads", -1))
                meta_by_file[f] = {"paired_end": paired, "paired_end_mode": paired, "read_length": 50,
                                   "norm_read_length": ext if ext > 0 else 50, "resolved_extend_bp": ext if ext > 0 else 0,
                                   "mapped_reads": n_reads, "norm_scale": 1.0e6 / (n_reads * (0.7 + 0.1 * k)), "threads": 1}
                files.append(f)
                for name in ("pos", "end", "flag", "mapq", "isize", "mate_same_tid"):
                    store[f"{tag}_s{k}_{name}"] = getattr(r, name)
                store[f"{tag}_s{k}_norm_scale"] = np.array(meta_by_file[f]["norm_scale"])
            intervals, matrix = rt.generate_chrom_matrix(CHROM, files, sizes, step, num_processors=1, **kw)
            store[f"{tag}_intervals"] = np.asarray(intervals, dtype=np.int64)
            store[f"{tag}_matrix"] = matrix
            store[f"{tag}_nsamples"] = np.array(len(samples))
            print(tag, matrix.shape, matrix.dtype, float(matrix.sum()))
    np.savez_compressed(OUT, **store)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB")


if __name__ == "__main__":
    main()

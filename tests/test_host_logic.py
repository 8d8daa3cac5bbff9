"""CPU-only checks of the host side: the C-ABI library loads and exports every declared symbol,
the numpy.sum restatement that defines the search bracket (dp.py:110-111) is bit-exact, layout and
budget-target rules, BED merge/sort semantics.  No kernel is launched here."""
import os
import re

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from rocco_b200 import _lib
    lib = _lib.load()                       # raises if the .so is missing or a symbol is absent
    header = open(os.path.join(REPO, "include", "rocco_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(rocco_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/rocco_b200.h but not exported"
    assert lib.rocco_b200_version().decode() == "0.1.0"


def test_compute_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import rocco_b200
    with pytest.raises(RuntimeError):
        rocco_b200.solve_penalized_chain(np.zeros(4), np.ones(3), 0.0)
    with pytest.raises(RuntimeError):
        rocco_b200.chrom_solution_to_bed("c", np.arange(0, 200, 50), np.array([1, 1, 0, 0], dtype=np.uint8))


def test_numpy_sum_restatement_is_bit_exact():
    from rocco_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)
    for n in [1, 2, 7, 8, 9, 127, 128, 129, 130, 255, 256, 1000, 4097, 65537, 934199, 1172352, 4979128]:
        a = rng.uniform(0, 3, size=n)
        assert lib.rocco_b200_numpy_sum_f64(a.ctypes.data, n) == float(np.sum(a))
        for g in (1.0, 6.86, 0.1, 2.5, 1e-3):
            assert _lib.numpy_sum_const(g, n) == float(np.sum(np.full(n, g))), (n, g)


def test_layout_and_budget_target():
    from rocco_b200.pipeline import layout_offsets, target_count_for_budget
    offs, total = layout_offsets([5, 16, 17, 1])
    assert offs == [0, 16, 32, 64] and total == 80
    # SURVEY.md appendix D item 12
    assert target_count_for_budget(934200, 0.02) == 18684
    assert target_count_for_budget(1172353, 0.045) == 52755
    assert target_count_for_budget(3120818, 0.015) == 46812
    assert target_count_for_budget(4979129, 0.03) == 149373


def test_merge_sorts_chromosomes_lexicographically(oracle):
    from rocco_b200.rocco import _merge_bed_records
    recs = [("chr2", 5, 10), ("chr10", 0, 5), ("chr1", 50, 60), ("chr1", 55, 70), ("chr1", 70, 80), ("chr10", 5, 7)]
    want = oracle.merge_bed_records(recs)
    assert _merge_bed_records(recs) == want == [("chr1", 50, 80), ("chr10", 0, 7), ("chr2", 5, 10)]
    assert _merge_bed_records(recs, min_length_bp=10) == oracle.merge_bed_records(recs, 10)


def test_combine_chrom_results_matches_oracle_on_messy_input(oracle, tmp_path):
    """vectorised combine == the reference's sort/merge semantics (rocco.py:74-95, 194-240): unsorted, overlapping,
    touching and nested records, extra columns, blank lines, lexicographic chromosome order, feature names"""
    from rocco_b200.rocco import combine_chrom_results
    rng = np.random.default_rng(0)
    files = []
    for k, chrom in enumerate(["chr2", "chr10", "chr1", "chrX"]):
        starts = rng.integers(0, 50_000, size=3000) * 10
        lens = rng.integers(1, 40, size=3000) * 10
        p = tmp_path / f"in{k}.bed"
        with open(p, "w") as fh:
            for i, (s, l) in enumerate(zip(starts, lens)):
                extra = "\tname\t0" if (k == 1 and i % 3 == 0) else ""
                fh.write(f"{chrom}\t{s}\t{s + l}{extra}\n")
                if i % 500 == 0:
                    fh.write("\n")
        files.append(str(p))
    (tmp_path / "empty.bed").write_text("")
    files.append(str(tmp_path / "empty.bed"))
    for names in (False, True):
        a = combine_chrom_results(files, str(tmp_path / "a.bed"), name_features=names)
        b = oracle.combine_chrom_results(files, str(tmp_path / "b.bed"), name_features=names)
        assert open(a).read() == open(b).read()
    with pytest.raises(FileNotFoundError):
        combine_chrom_results([str(tmp_path / "missing.bed")], str(tmp_path / "c.bed"))
    (tmp_path / "bad.bed").write_text("chr1\t5\n")
    with pytest.raises(ValueError):
        combine_chrom_results([str(tmp_path / "bad.bed")], str(tmp_path / "c.bed"))


def test_combine_native_and_fallback_paths_agree_with_oracle(oracle, tmp_path):
    """The native one-pass combine takes canonical BED text only; anything else (CRLF, padded lines, signed or
    underscore integers, whitespace-only lines) goes through the line reader -- both must give the oracle's file."""
    from rocco_b200 import _lib
    from rocco_b200.rocco import combine_chrom_results
    canon = tmp_path / "canon.bed"
    canon.write_text("chrB\t10\t20\nchrA\t5\t9\nchrA\t9\t12\n\nchrB\t15\t40\textra\nchr10\t007\t8\n")
    assert _lib.combine_bed_files([str(canon)], str(tmp_path / "n.bed")) == (3, True)
    assert open(tmp_path / "n.bed").read() == "chr10\t7\t8\nchrA\t5\t12\nchrB\t10\t40\n"
    messy = {
        "crlf.bed": "chr1\t1\t5\r\nchr1\t4\t9\r\n",
        "padded.bed": "  chr1\t100\t200  \n\t\nchr1\t150\t250\n",
        "signed.bed": "chr2\t+3\t1_0\nchr2\t-5\t2\n",
        "spaces_only.bed": "chr3\t1\t2\n   \nchr3\t2\t3\n",
    }
    for name, text in messy.items():
        p = tmp_path / name
        with open(p, "w", newline="") as fh:
            fh.write(text)
        assert _lib.combine_bed_files([str(p)], str(tmp_path / "x.bed")) is None, name
    files = [str(canon)] + [str(tmp_path / k) for k in messy]
    a = combine_chrom_results(files, str(tmp_path / "a.bed"))
    b = oracle.combine_chrom_results(files, str(tmp_path / "b.bed"))
    assert open(a).read() == open(b).read()
    a = combine_chrom_results([str(canon)], str(tmp_path / "a2.bed"), name_features=True)
    b = oracle.combine_chrom_results([str(canon)], str(tmp_path / "b2.bed"), name_features=True)
    assert open(a).read() == open(b).read()
    assert open(combine_chrom_results([], str(tmp_path / "none.bed"))).read() == ""


def test_forked_child_gets_a_named_error(monkeypatch):
    """A CUDA context does not survive fork(); the reference's callers fork worker pools (rocco.py:1176-1180)."""
    from rocco_b200 import _lib
    monkeypatch.setattr(_lib, "_CUDA_PID", os.getpid() + 1)          # as if the context had been created by a parent
    with pytest.raises(RuntimeError, match="forked child"):
        _lib.require_device()


def test_reorder_runs_equals_a_lexsort_of_the_records():
    from rocco_b200 import pipeline
    """the genome BED's record order (chromosome strings ascending, then start) from a block permutation of the shard's runs"""
    rng = np.random.default_rng(3)
    names = ["chr1", "chr2", "chr10", "chrX", "chr21", "chr3"]
    chrom, starts = [], []
    for k in range(len(names)):
        cnt = 0 if k == 2 else int(rng.integers(0, 40))
        chrom += [k] * cnt
        starts += sorted(rng.choice(10000, cnt, replace=False).tolist())
    runs = (np.array(chrom, np.int32), np.array(starts, np.int64), np.array(starts, np.int64) + 3)
    lex_order = sorted(range(len(names)), key=lambda k: names[k])
    got = pipeline.reorder_runs(runs, lex_order)
    lex_rank = np.array([sorted(names).index(c) for c in names])
    o = np.lexsort((runs[1], lex_rank[runs[0]]))
    assert np.array_equal(got[0], lex_rank[runs[0]][o]) and np.array_equal(got[1], runs[1][o]) and np.array_equal(got[2], runs[2][o])
    empty = pipeline.reorder_runs((np.zeros(0, np.int32), np.zeros(0, np.int64), np.zeros(0, np.int64)), lex_order)
    assert all(a.shape == (0,) for a in empty)


def test_bed_text_sizes_and_positioned_writes_reproduce_the_plain_writer(tmp_path):
    """the byte count announced per chromosome (digits of both coordinates, name, two tabs, newline) is exactly the text the
    native formatter produces, also at digit-count boundaries; positioned writes of the parts rebuild the one-call file"""
    from rocco_b200 import pipeline
    names = ["chr1", "chr10", "chrUn_KI270742v1", "chrX"]
    edge = np.array([0, 9, 10, 99, 100, 999, 1000, 99999, 100000, 4294967295, 4294967296, 10 ** 12], dtype=np.int64)
    rng = np.random.default_rng(5)
    chrom, starts = [], []
    for k in range(len(names)):
        vals = edge if k == 0 else (np.zeros(0, np.int64) if k == 2 else np.sort(rng.choice(5_000_000, 300, replace=False)))
        chrom += [k] * len(vals)
        starts += vals.tolist()
    runs = (np.array(chrom, np.int32), np.array(starts, np.int64), np.array(starts, np.int64) + 7)
    want = pipeline.runs_to_bed_text(names, runs, 1)
    sizes = pipeline.bed_text_sizes(names, runs, 1)
    per = [sum(len(line) + 1 for line in want.splitlines() if line.split("\t")[0] == nm) for nm in names]
    assert sizes.tolist() == per and sizes[2] == 0
    one = tmp_path / "one.bed"
    pipeline.runs_to_bed_file(str(one), names, runs, 1)
    assert one.read_text() == want
    parts = tmp_path / "parts.bed"
    parts.write_bytes(b"?" * (int(sizes.sum()) + 1000))                      # stale, longer content
    offsets = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    assert pipeline.write_genome_bed_part(str(parts), names, runs, 1, offsets) == int(sizes.sum())
    with open(parts, "ab") as fh:
        fh.truncate(int(sizes.sum()))
    assert parts.read_text() == want

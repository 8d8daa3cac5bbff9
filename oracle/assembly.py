"""ORACLE -- test infrastructure only.

CPU restatement of the step BEFORE the hot path: decoded alignment records ->
per-bin coverage -> scaled / trimmed / rounded track -> samples x bins count matrix.

  ccounts_backend.c:2416-2574   per-read filters, fragment inference, delta buffer + prefix sum   oracle/c/oracle_counts.c
  readtracks.py:468-518         count window, scaling, trim to the positive range, np.round        track_from_counts()
  readtracks.py:590-633         union of the samples' interval grids, scatter into the matrix      assemble_matrix()

Parity pin: readtracks.py's two ranges are pinned by running the REAL reference Python over this module's counter
(tests/golden/make_golden_assembly.py -> tests/golden/reference_assembly_v1_11_0.npz); the counting loop itself cannot
be pinned against the compiled reference (htslib is not buildable here): "parity unpinned" for those lines.
Only tests/ may import this module.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass

import numpy as np

from . import oracle as _orc

_I64 = ctypes.POINTER(ctypes.c_int64)


class CountOptions(ctypes.Structure):
    _fields_ = [("flag_include", ctypes.c_int32), ("flag_exclude", ctypes.c_int32), ("min_mapping_quality", ctypes.c_int32),
                ("paired_end_mode", ctypes.c_int32), ("one_read_per_bin", ctypes.c_int32),
                ("read_length", ctypes.c_int64), ("min_template_length", ctypes.c_int64), ("max_insert_size", ctypes.c_int64),
                ("shift_forward", ctypes.c_int64), ("shift_reverse", ctypes.c_int64), ("extend_bp", ctypes.c_int64)]


@dataclass
class Reads:
    """what htslib decodes per record: bam1_core_t pos / bam_endpos / flag / qual / isize and (mtid == tid)"""
    pos: np.ndarray
    end: np.ndarray
    flag: np.ndarray
    mapq: np.ndarray
    isize: np.ndarray
    mate_same_tid: np.ndarray


def synthetic_reads(n_reads: int, chrom_size: int, seed: int, read_length: int = 50, paired: bool = False) -> Reads:
    rng = np.random.default_rng(seed)
    # clustered starts (peaks) over a uniform background
    n_peak = n_reads // 3
    centers = rng.integers(0, chrom_size, size=max(1, n_reads // 400))
    pos = np.concatenate([rng.integers(0, max(1, chrom_size - read_length), size=n_reads - n_peak),
                          np.clip(rng.choice(centers, size=n_peak) + rng.normal(0, 150, size=n_peak).astype(np.int64), 0, chrom_size - 1)])
    pos = np.sort(pos).astype(np.int64)
    length = np.clip(read_length + rng.integers(-5, 6, size=n_reads), 20, None)
    flag = np.zeros(n_reads, dtype=np.uint16)
    flag[rng.random(n_reads) < 0.5] |= 16                                  # reverse strand
    flag[rng.random(n_reads) < 0.02] |= 1024                               # duplicates (excluded by the default 3844)
    flag[rng.random(n_reads) < 0.01] |= 256                                # secondary
    isize = np.zeros(n_reads, dtype=np.int64)
    same = np.ones(n_reads, dtype=np.uint8)
    if paired:
        flag |= 1
        proper = rng.random(n_reads) < 0.9
        flag[proper] |= 2
        first = rng.random(n_reads) < 0.5
        flag[first] |= 64
        flag[~first] |= 128
        frag = rng.integers(80, 600, size=n_reads)
        isize = np.where((flag & 16) == 0, frag, -frag).astype(np.int64)
        isize[rng.random(n_reads) < 0.01] = 0
        same = (rng.random(n_reads) < 0.98).astype(np.uint8)
        flag[rng.random(n_reads) < 0.01] |= 8
    mapq = rng.choice(np.array([0, 3, 10, 30, 42, 60], dtype=np.uint8), size=n_reads, p=[0.03, 0.03, 0.04, 0.2, 0.3, 0.4])
    return Reads(pos, (pos + length).astype(np.int64), flag, mapq, isize, same)


def count_alignment_region(reads: Reads, start: int, end: int, step: int, opt: CountOptions) -> np.ndarray:
    """ccounts_backend.c:2416-2574 over decoded records; returns float32 counts, one per `step` bp of [start, end)"""
    lib = ctypes.CDLL(_orc._PORT_LIB) if os.path.exists(_orc._PORT_LIB) else None
    if lib is None:
        _orc.build(ref=False)
        lib = ctypes.CDLL(_orc._PORT_LIB)
    n_bins = (int(end) - int(start) + int(step) - 1) // int(step)
    out = np.zeros(n_bins, dtype=np.float32)
    pos = np.ascontiguousarray(reads.pos, dtype=np.int64)
    endp = np.ascontiguousarray(reads.end, dtype=np.int64)
    flag = np.ascontiguousarray(reads.flag, dtype=np.uint16)
    mapq = np.ascontiguousarray(reads.mapq, dtype=np.uint8)
    isz = np.ascontiguousarray(reads.isize, dtype=np.int64)
    same = np.ascontiguousarray(reads.mate_same_tid, dtype=np.uint8)
    fn = lib.oracle_count_alignment_region
    fn.restype = ctypes.c_int
    st = fn(pos.ctypes.data_as(_I64), endp.ctypes.data_as(_I64), flag.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)),
            mapq.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), isz.ctypes.data_as(_I64),
            same.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), ctypes.c_size_t(pos.size), ctypes.byref(opt),
            ctypes.c_int64(int(start)), ctypes.c_int64(int(end)), ctypes.c_int64(int(step)),
            out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), ctypes.c_size_t(n_bins))
    if st != 0:
        raise MemoryError()
    return out


def count_window(read_min: int, read_max: int, chrom_size: int, step: int) -> tuple[int, int]:
    """readtracks.py:468-474: the counted window, snapped outwards to multiples of `step`"""
    count_start = max(0, (int(read_min) // step) * step)
    count_end = min(chrom_size, int(np.ceil(max(int(read_max), count_start + 1) / float(step)) * step))
    if count_end <= count_start:
        count_end = min(chrom_size, count_start + step)
    return count_start, count_end


def track_from_counts(counts, count_start: int, step: int, norm_scale: float, scale_by_step: bool = False,
                      const_scale: float = 1.0, round_digits: int = 5):
    """readtracks.py:492-518: scale, trim to [first positive, last positive], round"""
    vals = np.asarray(counts, dtype=np.float64)
    intervals = count_start + (np.arange(vals.size, dtype=np.int64) * int(step))
    vals = vals * float(norm_scale)
    if scale_by_step:
        vals = vals / float(step)
    if const_scale >= 0:
        vals = vals * const_scale
    positive = np.flatnonzero(vals > 0.0)
    if positive.size == 0:
        return None, None
    a, b = int(positive[0]), int(positive[-1]) + 1
    return intervals[a:b].astype(int), np.round(vals[a:b], round_digits)


def assemble_matrix(tracks, low_memory: bool = False):
    """readtracks.py:590-633: samples without data are dropped, the columns are the sorted union of the samples'
    interval starts, every sample's values are scattered to their columns, the rest stays zero"""
    kept = [(iv, v) for iv, v in tracks if iv is not None and v is not None]
    if not kept:
        return None, None
    common = np.sort(np.unique(np.concatenate([iv for iv, _ in kept], axis=0)))
    dtype = np.float32 if low_memory else np.float64
    matrix = np.zeros((len(kept), len(common)), dtype=dtype)
    for i, (iv, v) in enumerate(kept):
        matrix[i, np.searchsorted(common, iv)] = np.asarray(v, dtype=dtype)
    return np.array(common).astype(int), matrix

"""Device-side matrix assembly: the step of the reference's ``rocco/readtracks.py`` that feeds the hot path
(SURVEY.md section 8(f) rank 3).

The reference counts every BAM into a float32 coverage vector on the host (``native/ccounts_backend.c:2416-2574``),
scales / trims / rounds it in NumPy (``readtracks.py:492-518``), builds the sorted union of the samples' interval
starts and scatters every sample into a float64 ``[samples, bins]`` matrix (``readtracks.py:590-633``) -- which the hot
path would then have to upload at 8 bytes per sample-bin.  Here the decoded alignment records (what htslib hands
over: position, end, flag, mapping quality, template length, mate-on-same-reference) are what crosses PCIe; the coverage,
the scaling / trimming / rounding and the union + scatter run on the GPU and the matrix is born in HBM in the layout
``score_loci_wls`` reads (``rocco_b200.pipeline.score_loci_wls_device`` takes it as is).

BAM / BGZF decoding itself is out of scope (it stays with htslib on the CPU); so are bigWig inputs.
There is no CPU fallback: without the CUDA library or a device these functions raise ``RuntimeError``.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import numpy as np

from . import _lib


def _torch():
    import torch
    return torch


def _stream_ptr(device) -> int:
    return int(_torch().cuda.current_stream(device).cuda_stream)


def count_window(read_min: int, read_max: int, chrom_size: int, step: int) -> tuple[int, int]:
    """The counted window of one sample (``readtracks.py:468-474``): [first mapped base, last mapped base) snapped
    outwards to multiples of ``step`` and clipped to the chromosome."""
    step = int(step)
    start = max(0, (int(read_min) // step) * step)
    end = min(int(chrom_size), int(np.ceil(max(int(read_max), start + 1) / float(step)) * step))
    if end <= start:
        end = min(int(chrom_size), start + step)
    return start, end


def count_alignment_region(pos, end, flag, mapq, isize, mate_same_tid, start: int, stop: int, step: int, read_length: int,
                           one_read_per_bin: int = 0, flag_include: int = 0, flag_exclude: int = 0, extend_bp: int = 0,
                           paired_end_mode: int = 0, min_mapping_quality: int = 0, min_template_length: int = -1,
                           max_insert_size: int = 1000, shift_forward_strand53: int = 0, shift_reverse_strand53: int = 0,
                           device=None):
    """Coverage of ``[start, stop)`` in bins of ``step`` bp from decoded alignment records -- the arguments after the record
    arrays are those of the reference's ``_hts_counts.count_alignment_region`` (``count_mode="coverage"``).

    The record arrays may be NumPy arrays (uploaded here) or CUDA tensors.  Returns a float32 CUDA tensor."""
    torch = _torch()
    lib = _lib.load()
    _lib.require_device()
    dev = device or torch.device("cuda", torch.cuda.current_device())

    def dev_array(a, dtype):
        if isinstance(a, torch.Tensor):
            return a.to(device=dev, dtype=dtype).contiguous()
        return torch.from_numpy(np.ascontiguousarray(a, dtype={torch.int64: np.int64, torch.int16: np.int16, torch.uint8: np.uint8}[dtype])).to(dev)

    n_reads = int(len(pos))
    d_pos, d_end, d_isize = dev_array(pos, torch.int64), dev_array(end, torch.int64), dev_array(isize, torch.int64)
    flag16 = np.ascontiguousarray(flag, dtype=np.uint16).view(np.int16) if not isinstance(flag, torch.Tensor) else flag
    d_flag = dev_array(flag16, torch.int16)
    d_mapq, d_same = dev_array(mapq, torch.uint8), dev_array(mate_same_tid, torch.uint8)
    if int(step) <= 0 or int(stop) <= int(start):
        raise ValueError("`step` must be positive and the region non-empty")
    n_bins = (int(stop) - int(start) + int(step) - 1) // int(step)
    counts = torch.empty(n_bins, dtype=torch.float32, device=dev)
    opt = _lib.CountOptions(flag_include=max(0, int(flag_include)), flag_exclude=max(0, int(flag_exclude)),
                            min_mapping_quality=int(min_mapping_quality), paired_end_mode=int(paired_end_mode),
                            one_read_per_bin=int(bool(one_read_per_bin)), read_length=int(read_length),
                            min_template_length=int(min_template_length), max_insert_size=int(max_insert_size),
                            shift_forward_strand53=int(shift_forward_strand53), shift_reverse_strand53=int(shift_reverse_strand53),
                            extend_bp=int(extend_bp))
    with torch.cuda.device(dev):
        st = lib.rocco_b200_count_alignment_region_dev(
            d_pos.data_ptr(), d_end.data_ptr(), d_flag.data_ptr(), d_mapq.data_ptr(), d_isize.data_ptr(), d_same.data_ptr(),
            n_reads, ctypes.byref(opt), int(start), int(stop), int(step), counts.data_ptr(), n_bins, ctypes.c_void_p(_stream_ptr(dev)))
    _lib.check(st, "count_alignment_region")
    return counts


def assemble_chrom_matrix(tracks: Sequence[dict], step: int, round_digits: int = 5, low_memory: bool = False):
    """``generate_chrom_matrix``'s tail (``readtracks.py:590-633``) with ``get_bam_chrom_reads``' post-processing
    (``492-518``) folded in, on the device.

    ``tracks``: one dict per sample with ``counts`` (float32 CUDA tensor, the counted window), ``count_start`` (bp, a
    multiple of ``step``), ``norm_scale`` and optionally ``const_scale`` (default 1.0) and ``scale_by_step`` (False).
    Samples whose scaled track has no positive value are excluded, like the reference excludes them.
    Returns ``(intervals, matrix)``: int64 NumPy interval starts and the ``[kept samples, bins]`` CUDA tensor (float64, or
    float32 with ``low_memory``); ``(None, None)`` when no sample has data."""
    torch = _torch()
    lib = _lib.load()
    _lib.require_device()
    if not tracks:
        return None, None
    step = int(step)
    dev = tracks[0]["counts"].device
    arr = (_lib.Track * len(tracks))()
    keep_alive = []
    for k, t in enumerate(tracks):
        c = t["counts"]
        if not (isinstance(c, torch.Tensor) and c.is_cuda and c.dtype == torch.float32 and c.is_contiguous()):
            raise ValueError("`counts` must be a contiguous float32 CUDA tensor")
        if int(t["count_start"]) % step != 0:
            raise ValueError("`count_start` must be a multiple of `step`")
        keep_alive.append(c)
        arr[k] = _lib.Track(d_counts=c.data_ptr(), count_len=int(c.shape[0]), count_start=int(t["count_start"]),
                            norm_scale=float(t["norm_scale"]), const_scale=float(t.get("const_scale", 1.0)),
                            scale_by_step=int(bool(t.get("scale_by_step", False))))
    first = np.empty(len(tracks), dtype=np.int64)
    last = np.empty(len(tracks), dtype=np.int64)
    with torch.cuda.device(dev):
        _lib.check(lib.rocco_b200_track_positive_range_dev(arr, len(tracks), step, _lib.np_ptr(first), _lib.np_ptr(last),
                                                           ctypes.c_void_p(_stream_ptr(dev))), "track ranges")
    kept = [k for k in range(len(tracks)) if first[k] >= 0]
    if not kept:
        return None, None
    # union of the kept samples' [first, last] bin ranges on the common step lattice = a few maximal runs of bins
    spans = sorted((int(tracks[k]["count_start"]) + int(first[k]) * step, int(tracks[k]["count_start"]) + (int(last[k]) + 1) * step) for k in kept)
    seg_start, seg_end = [spans[0][0]], [spans[0][1]]
    for a, b in spans[1:]:
        if a <= seg_end[-1]:
            seg_end[-1] = max(seg_end[-1], b)
        else:
            seg_start.append(a)
            seg_end.append(b)
    seg_cols = np.array([(b - a) // step for a, b in zip(seg_start, seg_end)], dtype=np.int64)
    seg_first_col = np.concatenate([[0], np.cumsum(seg_cols)[:-1]]).astype(np.int64)
    n_cols = int(seg_cols.sum())
    seg_start_np = np.array(seg_start, dtype=np.int64)
    matrix = torch.empty((len(kept), n_cols), dtype=torch.float32 if low_memory else torch.float64, device=dev)
    d_intervals = torch.empty(n_cols, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        st = lib.rocco_b200_assemble_matrix_dev(arr, _lib.np_ptr(first), _lib.np_ptr(last), len(tracks), step, int(round_digits),
                                                _lib.np_ptr(seg_start_np), _lib.np_ptr(seg_first_col), len(seg_start), n_cols,
                                                1 if low_memory else 0, matrix.data_ptr(), d_intervals.data_ptr(),
                                                ctypes.c_void_p(_stream_ptr(dev)))
    _lib.check(st, "assemble_chrom_matrix")
    del keep_alive
    return d_intervals.cpu().numpy().astype(int), matrix


def generate_chrom_matrix_from_reads(samples: Sequence[dict], chrom_size: int, step: int, const_scale: float = 1.0,
                                     round_digits: int = 5, scale_by_step: bool = False, min_mapping_score: int = 10,
                                     flag_include: Optional[int] = None, flag_exclude: int = 3844, center_reads: bool = False,
                                     low_memory: bool = False, device=None):
    """``generate_chrom_matrix`` (``readtracks.py:521-633``) for BAM inputs whose records are already decoded.

    ``samples``: one dict per BAM with the record arrays ``pos, end, flag, mapq, isize, mate_same_tid`` and the per-file
    metadata the reference derives in ``_get_bam_count_metadata``: ``read_length``, ``resolved_extend_bp``,
    ``paired_end_mode``, ``norm_scale``.  The keyword arguments keep the reference's names and defaults.
    Returns ``(intervals, matrix)`` like ``assemble_chrom_matrix``."""
    tracks = []
    for s in samples:
        flag = np.asarray(s["flag"])
        pos, end = np.asarray(s["pos"]), np.asarray(s["end"])
        mapped = (flag & max(0, int(flag_exclude))) == 0              # get_alignment_chrom_range: records passing the exclude mask
        if not mapped.any():
            continue
        start, stop = count_window(int(pos[mapped].min()), int(end[mapped].max()), chrom_size, step)
        counts = count_alignment_region(pos, end, flag, s["mapq"], s["isize"], s["mate_same_tid"], start, stop, step,
                                        int(s["read_length"]), one_read_per_bin=1 if center_reads else 0,
                                        flag_include=max(0, int(flag_include or 0)), flag_exclude=max(0, int(flag_exclude)),
                                        extend_bp=max(0, int(s.get("resolved_extend_bp", 0))),
                                        paired_end_mode=1 if bool(s.get("paired_end_mode", False)) else 0,
                                        min_mapping_quality=max(0, int(min_mapping_score)), device=device)
        tracks.append({"counts": counts, "count_start": start, "norm_scale": float(s["norm_scale"]),
                       "const_scale": float(const_scale), "scale_by_step": bool(scale_by_step)})
    if not tracks:
        return None, None
    return assemble_chrom_matrix(tracks, step, round_digits=round_digits, low_memory=low_memory)

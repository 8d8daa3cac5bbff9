"""Device-resident driver of the hot path for MANY chromosomes at once.

The reference walks chromosomes one by one (rocco.py:948) and solves them in a fork pool of <= 4
processes (rocco.py:1146-1184).  Here every chromosome of a rank's shard is laid out in one device
buffer and each stage is a single batched launch set: scores [sum n_c] float64, masks [sum n_c]
uint8, chromosome c at ``offset_c`` (offsets padded to 16 elements so 128-bit accesses stay aligned).
PyTorch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import numpy as np

from . import _lib

ALIGN = 16


def _torch():
    import torch
    return torch


def _stream_ptr(device) -> int:
    torch = _torch()
    return int(torch.cuda.current_stream(device).cuda_stream)


def layout_offsets(lengths: Sequence[int]) -> tuple[list[int], int]:
    offsets, cur = [], 0
    for n in lengths:
        offsets.append(cur)
        cur += (int(n) + ALIGN - 1) // ALIGN * ALIGN
    return offsets, max(cur, ALIGN)


def target_count_for_budget(n: int, budget: float) -> int:
    """dp.py:197 -- computed on the host in float64 exactly as the reference does."""
    return int(np.floor(int(n) * float(budget)))


def pack_scores(scores_list, device=None):
    """Concatenate per-chromosome score vectors (NumPy or CUDA tensors) into one float64 device buffer."""
    torch = _torch()
    if device is None:
        device = next((s.device for s in scores_list if isinstance(s, torch.Tensor) and s.is_cuda),
                      torch.device("cuda", torch.cuda.current_device()))
    lengths = [int(s.shape[0]) for s in scores_list]
    offsets, total = layout_offsets(lengths)
    buf = torch.zeros(total, dtype=torch.float64, device=device)
    for s, off, n in zip(scores_list, offsets, lengths):
        if isinstance(s, torch.Tensor):
            if s.dim() != 1:
                raise ValueError("`scores` must be a one-dimensional array")
            buf[off:off + n].copy_(s.to(dtype=torch.float64), non_blocking=True)
        else:
            a = np.ascontiguousarray(s, dtype=np.float64)
            if a.ndim != 1:
                raise ValueError("`scores` must be a one-dimensional array")
            buf[off:off + n].copy_(torch.from_numpy(a), non_blocking=False)
    return buf, offsets, lengths


def solve_packed(d_scores, offsets, lengths, budgets, gammas, selection_penalties=None, max_iter: int = 60,
                 levels_per_round: int = 0, d_masks=None):
    """Run the batched chain solve on an already packed device buffer.

    Returns (d_masks uint8 tensor, list of ChainResult-like dicts).  No host copy of the masks."""
    torch = _torch()
    lib = _lib.load()
    _lib.require_device()
    nchrom = len(lengths)
    if selection_penalties is None:
        selection_penalties = [None] * nchrom
    tasks = (_lib.ChainTask * nchrom)()
    for c in range(nchrom):
        n = int(lengths[c])
        if n <= 0:
            raise ValueError("`scores` cannot be empty")
        g = float(gammas[c])
        t = tasks[c]
        t.offset, t.n, t.gamma = int(offsets[c]), n, g
        t.cost_sum = _lib.numpy_sum_const(g, n - 1) if n > 1 else 0.0
        t.max_iter = int(max_iter)
        if selection_penalties[c] is not None:
            t.mode, t.selection_penalty, t.target_count = 0, float(selection_penalties[c]), 0
        elif budgets[c] is None:
            t.mode, t.selection_penalty, t.target_count = 0, 0.0, 0
        else:
            t.mode, t.selection_penalty = 1, 0.0
            t.target_count = target_count_for_budget(n, budgets[c])
    if d_masks is None:
        d_masks = torch.zeros(d_scores.shape[0], dtype=torch.uint8, device=d_scores.device)
    results = (_lib.ChainResult * nchrom)()
    with torch.cuda.device(d_scores.device):
        st = lib.rocco_b200_chain_solve_batch_dev(
            ctypes.c_void_p(d_scores.data_ptr()), None, tasks, nchrom,
            ctypes.c_void_p(d_masks.data_ptr()), results, int(levels_per_round),
            ctypes.c_void_p(_stream_ptr(d_scores.device)))
    _lib.check(st, "solve_chrom_exact")
    out = []
    for c in range(nchrom):
        r = results[c]
        out.append({
            "selection_penalty": float(r.selection_penalty),
            "penalized_objective": float(r.penalized_objective),
            "objective": float(r.objective),
            "selected_count": int(r.selected_count),
            "switch_count": int(r.switch_count),
            "exact_tie_bins": int(r.exact_tie_bins),
            "near_tie_bins": int(r.near_tie_bins),
            "dp_passes": int(r.dp_passes),
            "search_rounds": int(r.search_rounds),
        })
    return d_masks, out


def solve_chromosomes(scores_list, budgets, gammas, selection_penalties=None, max_iter: int = 60,
                      levels_per_round: int = 0):
    """solve_chrom_exact for a list of chromosomes in one batched launch set; NumPy masks returned."""
    d_scores, offsets, lengths = pack_scores(scores_list)
    d_masks, res = solve_packed(d_scores, offsets, lengths, budgets, gammas, selection_penalties,
                                max_iter=max_iter, levels_per_round=levels_per_round)
    h_masks = d_masks.cpu().numpy()
    for r, off, n in zip(res, offsets, lengths):
        r["solution"] = h_masks[off:off + n].copy()
    return res


def masks_to_runs(d_masks, offsets, lengths):
    """All maximal runs of selected bins (last bin of each chromosome dropped), one batched pass.

    Returns (chrom_index int32[k], start_bin int64[k], end_bin int64[k]) with end exclusive."""
    torch = _torch()
    lib = _lib.load()
    nchrom = len(lengths)
    cap = int(sum((int(n) + 1) // 2 for n in lengths)) + 1
    starts = np.empty(cap, dtype=np.int64)
    ends = np.empty(cap, dtype=np.int64)
    chrom = np.empty(cap, dtype=np.int32)
    offs = np.asarray(offsets, dtype=np.uint64)
    lens = np.asarray(lengths, dtype=np.uint64)
    with torch.cuda.device(d_masks.device):
        k = lib.rocco_b200_mask_to_runs_batch_dev(
            ctypes.c_void_p(d_masks.data_ptr()), _lib.np_ptr(offs), _lib.np_ptr(lens), nchrom,
            _lib.np_ptr(starts), _lib.np_ptr(ends), _lib.np_ptr(chrom), cap,
            ctypes.c_void_p(_stream_ptr(d_masks.device)))
    if k < 0:
        _lib.check(int(k), "mask_to_runs")
    return chrom[:k].copy(), starts[:k].copy(), ends[:k].copy()

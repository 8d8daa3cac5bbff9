"""Device-resident driver of the hot path for MANY chromosomes at once.

The reference walks chromosomes one by one (rocco.py:948) and solves them in a fork pool of <= 4
processes (rocco.py:1146-1184).  Here every chromosome of a rank's shard is laid out in one device
buffer and each stage is a single batched launch set: scores [sum n_c] float64, masks [sum n_c]
uint8, chromosome c at ``offset_c`` (offsets padded to 16 elements so 128-bit accesses stay aligned).
PyTorch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes
import threading
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
    torch = _torch()
    if any(isinstance(s, torch.Tensor) and s.is_cuda for s in scores_list):
        d_scores, offsets, lengths = pack_scores(scores_list)
        d_masks, res = solve_packed(d_scores, offsets, lengths, budgets, gammas, selection_penalties,
                                    max_iter=max_iter, levels_per_round=levels_per_round)
        h_masks = d_masks.cpu().numpy()
    else:
        # Host inputs: run on this thread's own high-priority stream with pinned staging.  The solve is a chain of short
        # launches; on the shared default stream (or at normal priority) it queues behind whatever bulk scoring kernels
        # and 256 MB copies other host threads have in flight, and a pageable upload is cut into many small DMA commands
        # each of which waits for one of those copies.
        _lib.require_device()
        dev = torch.device("cuda", torch.cuda.current_device())
        arrays = []
        for s in scores_list:
            a = np.ascontiguousarray(s, dtype=np.float64)
            if a.ndim != 1:
                raise ValueError("`scores` must be a one-dimensional array")
            arrays.append(a)
        lengths = [int(a.shape[0]) for a in arrays]
        offsets, total = layout_offsets(lengths)
        with _StageLease(total * 9) as stage, _UrgentStream(dev) as stream:
            h_scores = stage[: total * 8].view(torch.float64)
            h_scores.zero_()
            for a, off, n in zip(arrays, offsets, lengths):
                h_scores[off:off + n].copy_(torch.from_numpy(a))
            with torch.cuda.stream(stream):
                d_scores = torch.empty(total, dtype=torch.float64, device=dev)
                # kernel-driven upload: a DMA command would wait for every bulk copy already queued (see runtime.cu)
                _lib.check(_lib.load().rocco_b200_pull_pinned(ctypes.c_void_p(d_scores.data_ptr()), ctypes.c_void_p(h_scores.data_ptr()),
                                                              total * 8, ctypes.c_void_p(stream.cuda_stream)), "upload")
                d_masks, res = solve_packed(d_scores, offsets, lengths, budgets, gammas, selection_penalties,
                                            max_iter=max_iter, levels_per_round=levels_per_round)
                h_m = stage[total * 8: total * 9]
                h_m.copy_(d_masks, non_blocking=True)
                stream.synchronize()
            h_masks = h_m.numpy().copy()
    for r, off, n in zip(res, offsets, lengths):
        r["solution"] = h_masks[off:off + n].copy()
    return res


_URGENT: dict = {}


class _UrgentStream:
    """Borrow a highest-priority stream of `dev` from a process-wide free list (returned on exit).  The set of streams
    stays as small as the peak number of concurrent host threads, so torch's per-stream caching allocator reaches a steady
    state instead of seeing a fresh stream -- and a fresh cudaMalloc -- for every short-lived worker thread."""

    def __init__(self, dev):
        self.key = (dev.type, dev.index)
        self.dev = dev
        self.stream = None

    def __enter__(self):
        torch = _torch()
        with _STAGES_LOCK:
            free = _URGENT.setdefault(self.key, [])
            self.stream = free.pop() if free else None
        if self.stream is None:
            self.stream = torch.cuda.Stream(device=self.dev, priority=-1)
        return self.stream

    def __exit__(self, *exc):
        with _STAGES_LOCK:
            _URGENT[self.key].append(self.stream)
        self.stream = None
        return False


_STAGES: list = []
_STAGES_LOCK = threading.Lock()


class _StageLease:
    """A pinned staging buffer (uint8) borrowed from a process-wide free list and returned on exit.  Buffers are sized in
    64 MB steps and never freed: allocating or freeing pinned memory synchronises the device, which would stall the bulk
    copies and kernels of every other host thread."""

    def __init__(self, nbytes: int):
        self.nbytes = int(nbytes)
        self.buf = None

    def __enter__(self):
        torch = _torch()
        with _STAGES_LOCK:
            fit = [k for k, b in enumerate(_STAGES) if b.numel() >= self.nbytes]
            if fit:
                self.buf = _STAGES.pop(min(fit, key=lambda k: _STAGES[k].numel()))
        if self.buf is None:
            step = 64 << 20
            self.buf = torch.empty(((self.nbytes + step - 1) // step) * step or step, dtype=torch.uint8, pin_memory=True)
        return self.buf

    def __exit__(self, *exc):
        with _STAGES_LOCK:
            _STAGES.append(self.buf)
        self.buf = None
        return False


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


# ---------------------------------------------------------------------------------------------
# scoring on device-resident matrices
# ---------------------------------------------------------------------------------------------
def score_params(lower_bound_z=1.0, prior_df=5.0, min_effect=None, precision_floor_ratio=0.01,
                 spatial_window=31, baseline_window=101, exact_pilot=False):
    lib = _lib.load()
    prm = _lib.ScoreParams()
    lib.rocco_b200_default_score_params(ctypes.byref(prm))
    prm.lower_bound_z = float(lower_bound_z)
    prm.prior_df = float(prior_df)
    prm.use_min_effect = 0 if min_effect is None else 1
    prm.min_effect = 0.0 if min_effect is None else max(float(min_effect), 0.0)
    prm.precision_floor_ratio = float(max(precision_floor_ratio, 0.0))
    prm.spatial_window = int(spatial_window)
    prm.baseline_window = int(baseline_window)
    prm.pilot_mode = 1 if exact_pilot else 0          # exact np.median row offsets for any row length (validation mode)
    return prm


def score_loci_wls_device(d_matrix, out_scores=None, params=None, details: bool = False):
    """score_loci_wls on a CUDA tensor [samples, bins] (float64 or float32, C-contiguous).

    Writes the scores into ``out_scores`` (a float64 CUDA tensor/view of length bins) or a new
    tensor; with ``details`` also returns the per-locus detail tensors.  Nothing leaves the device."""
    torch = _torch()
    lib = _lib.load()
    if d_matrix.dim() != 2 or not d_matrix.is_cuda or not d_matrix.is_contiguous():
        raise ValueError("`d_matrix` must be a contiguous two-dimensional CUDA tensor")
    if d_matrix.dtype not in (torch.float64, torch.float32):
        raise ValueError("`d_matrix` must be float64 or float32")
    m, n = d_matrix.shape
    if m == 0 or n == 0:
        raise ValueError("`chrom_matrix` must be non-empty")
    dev = d_matrix.device
    if out_scores is None:
        out_scores = torch.empty(n, dtype=torch.float64, device=dev)
    out = _lib.ScoreOutputs()
    out.scores = out_scores.data_ptr()
    extra = {}
    if details:
        for k in ("mean", "raw_variance", "prior_variance", "moderated_variance", "standard_error"):
            extra[k] = torch.empty(n, dtype=torch.float64, device=dev)
            setattr(out, k, extra[k].data_ptr())
        extra["centered_matrix"] = torch.empty((m, n), dtype=torch.float64, device=dev)
        out.centered_matrix = extra["centered_matrix"].data_ptr()
    prm = params if params is not None else score_params()
    with torch.cuda.device(dev):
        st = lib.rocco_b200_score_loci_wls_dev(
            ctypes.c_void_p(d_matrix.data_ptr()), 1 if d_matrix.dtype == torch.float32 else 0, m, n,
            ctypes.byref(prm), ctypes.byref(out), ctypes.c_void_p(_stream_ptr(dev)))
    if st == _lib.ST_NONFINITE:
        raise ValueError("Locus scoring produced non-finite values (or `chrom_matrix` contains non-finite values)")
    _lib.check(st, "score_loci_wls")
    if not details:
        return out_scores
    extra["total_df"] = float(out.total_df)
    extra["prior_spatial_window"] = int(out.resolved_spatial_window)
    extra["local_baseline_window"] = int(out.baseline_window)
    extra["local_baseline_lambda"] = float(out.baseline_lambda)
    return out_scores, extra


def lpt_partition(weights: Sequence[int], parts: int) -> list[list[int]]:
    """Longest-processing-time greedy packing of chromosomes onto ranks by bin count."""
    order = sorted(range(len(weights)), key=lambda i: (-int(weights[i]), i))
    loads = [0] * parts
    out: list[list[int]] = [[] for _ in range(parts)]
    for i in order:
        r = min(range(parts), key=lambda k: (loads[k], k))
        out[r].append(i)
        loads[r] += int(weights[i])
    for r in range(parts):
        out[r].sort()
    return out


_SCORE_STREAMS: dict = {}


def _score_streams(dev, count: int):
    torch = _torch()
    key = (dev.index, count)
    if key not in _SCORE_STREAMS:
        _SCORE_STREAMS[key] = [torch.cuda.Stream(device=dev) for _ in range(count)]
    return _SCORE_STREAMS[key]


def run_shard(d_matrices, budgets, gammas, params=None, levels_per_round: int = 0, want_runs: bool = True,
              score_streams: int = 1):
    """The hot path for one rank's chromosomes, device-resident end to end:
    score every chromosome -> one batched budget search + solve -> one batched mask->runs pass.

    With `score_streams` > 1 chromosomes are scored from that many host threads, each on its own CUDA stream
    (measured on B200: no gain at 100 samples -- the streaming kernels already fill the device -- so the default is 1).
    Returns dict(d_scores, d_masks, offsets, lengths, results, runs)."""
    torch = _torch()
    lengths = [int(x.shape[1]) for x in d_matrices]
    offsets, total = layout_offsets(lengths)
    dev = d_matrices[0].device
    d_scores = torch.zeros(total, dtype=torch.float64, device=dev)
    if score_streams <= 1 or len(d_matrices) == 1:
        for x, off, n in zip(d_matrices, offsets, lengths):
            score_loci_wls_device(x, out_scores=d_scores[off:off + n], params=params)
    else:
        from concurrent.futures import ThreadPoolExecutor
        streams = _score_streams(dev, score_streams)
        main = torch.cuda.current_stream(dev)
        ready = torch.cuda.Event()
        ready.record(main)
        # largest chromosomes first, alternating streams
        order = sorted(range(len(d_matrices)), key=lambda i: -lengths[i])

        def work(slot):
            st = streams[slot]
            st.wait_event(ready)
            with torch.cuda.device(dev), torch.cuda.stream(st):
                for i in order[slot::score_streams]:
                    score_loci_wls_device(d_matrices[i], out_scores=d_scores[offsets[i]:offsets[i] + lengths[i]], params=params)
            ev = torch.cuda.Event()
            ev.record(st)
            return ev

        with ThreadPoolExecutor(max_workers=score_streams) as pool:
            for ev in pool.map(work, range(score_streams)):
                main.wait_event(ev)
    d_masks, results = solve_packed(d_scores, offsets, lengths, budgets, gammas, levels_per_round=levels_per_round)
    runs = masks_to_runs(d_masks, offsets, lengths) if want_runs else None
    return {"d_scores": d_scores, "d_masks": d_masks, "offsets": offsets, "lengths": lengths,
            "results": results, "runs": runs}


def runs_to_bed_file(path: str, chrom_names, runs, step: int, first_start: int = 0) -> str:
    """Write all runs as BED3 (chromosomes in the given order; coordinates = first_start + bin*step) with the native
    formatter -- no per-record Python work."""
    chrom, starts, ends = runs
    return _lib.write_bed_arrays(path, list(chrom_names), chrom, first_start + starts * step, first_start + ends * step)


_POW10 = np.array([10 ** k for k in range(1, 19)], dtype=np.int64)


def bed_text_sizes(chrom_names, runs, step: int, first_start: int = 0) -> np.ndarray:
    """Bytes of BED3 text each chromosome of `runs` will take (name, two tabs, newline, the digits of both coordinates) --
    what a rank announces to the others before a genome BED is assembled with positioned writes (`write_genome_bed_part`)."""
    chrom, starts, ends = runs
    k = len(chrom_names)
    if chrom.shape[0] == 0:
        return np.zeros(k, dtype=np.int64)
    s = first_start + starts * step
    e = first_start + ends * step
    if int(s.min(initial=0)) < 0:
        raise ValueError("negative coordinates")
    digits = (np.searchsorted(_POW10, s, side="right") + 1) + (np.searchsorted(_POW10, e, side="right") + 1)
    name_len = np.array([len(str(c).encode("utf-8")) for c in chrom_names], dtype=np.int64)
    per_chrom_digits = np.bincount(chrom, weights=digits, minlength=k).astype(np.int64)
    counts = np.bincount(chrom, minlength=k).astype(np.int64)
    return per_chrom_digits + counts * (name_len + 3)


def write_genome_bed_part(path: str, chrom_names, runs, step: int, offsets, first_start: int = 0) -> int:
    """This rank's chromosomes written into the shared genome BED at their byte offsets (`offsets[c]` for chromosome index c
    of `chrom_names`; chromosomes without records are skipped).  Returns the bytes written."""
    chrom, starts, ends = runs
    bounds = np.searchsorted(chrom, np.arange(len(chrom_names) + 1))
    total = 0
    for c, name in enumerate(chrom_names):
        a, b = int(bounds[c]), int(bounds[c + 1])
        if b > a:
            total += _lib.write_bed_arrays_at(path, int(offsets[c]), [name], None, first_start + starts[a:b] * step,
                                              first_start + ends[a:b] * step)
    return total


def reorder_runs(runs, chrom_order):
    """Runs regrouped so that chromosome index `chrom_order[0]` comes first, then `chrom_order[1]`, ...  The runs of a
    shard arrive grouped by chromosome and sorted by start inside each group (`mask_to_runs`), so putting a genome into
    the reference's record order (rocco.py:74-95: chromosome STRINGS ascending, then start) is a permutation of
    blocks -- no sort.  Returns (position in chrom_order as int32, starts, ends)."""
    chrom, starts, ends = runs
    order = np.asarray(chrom_order, dtype=np.int64)
    bounds = np.searchsorted(chrom, np.arange(int(order.max(initial=-1)) + 2))
    take = np.concatenate([np.arange(bounds[c], bounds[c + 1]) for c in order] or [np.zeros(0, np.int64)])
    sizes = bounds[order + 1] - bounds[order]
    return np.repeat(np.arange(order.shape[0], dtype=np.int32), sizes), starts[take], ends[take]


def runs_to_bed_text(chrom_names, runs, step: int, first_start: int = 0) -> str:
    """BED3 text of all runs (per chromosome in the given order; coordinates = first_start + bin*step)."""
    chrom, starts, ends = runs
    s = first_start + starts * step
    e = first_start + ends * step
    names = np.asarray(chrom_names, dtype=object)[chrom]
    return "".join(f"{c}\t{a}\t{b}\n" for c, a, b in zip(names.tolist(), s.tolist(), e.tolist()))


def sweep_multipliers(scores, gamma: float, lambdas):
    """Selected count, penalized objective and objective of one chromosome for MANY multipliers
    (BASELINE.json config 5: a 256-multiplier budget sweep) -- up to 256 per launch set, the scores
    read once per tile whatever the number of multipliers.  Returns three NumPy arrays."""
    torch = _torch()
    lib = _lib.load()
    _lib.require_device()
    if isinstance(scores, torch.Tensor):
        d_scores = scores.to(dtype=torch.float64).contiguous()
        if not d_scores.is_cuda:
            d_scores = d_scores.cuda()
    else:
        d_scores = torch.from_numpy(np.ascontiguousarray(scores, dtype=np.float64)).cuda()
    if d_scores.dim() != 1 or d_scores.shape[0] == 0:
        raise ValueError("`scores` must be a non-empty one-dimensional array")
    lam = np.ascontiguousarray(lambdas, dtype=np.float64)
    k = lam.shape[0]
    counts = np.zeros(k, dtype=np.int64)
    pen = np.zeros(k, dtype=np.float64)
    obj = np.zeros(k, dtype=np.float64)
    with torch.cuda.device(d_scores.device):
        st = lib.rocco_b200_chain_sweep_dev(
            ctypes.c_void_p(d_scores.data_ptr()), int(d_scores.shape[0]), float(gamma), _lib.np_ptr(lam), int(k),
            _lib.np_ptr(counts), _lib.np_ptr(pen), _lib.np_ptr(obj), ctypes.c_void_p(_stream_ptr(d_scores.device)))
    _lib.check(st, "sweep_multipliers")
    return counts, pen, obj


def score_partial_device(d_matrix_local, params=None):
    """Per-sample stages of this rank's rows -> the four per-bin accumulators, a [4, bins] float64 CUDA tensor."""
    torch = _torch()
    lib = _lib.load()
    if d_matrix_local.dim() != 2 or not d_matrix_local.is_cuda or not d_matrix_local.is_contiguous():
        raise ValueError("`d_matrix_local` must be a contiguous two-dimensional CUDA tensor")
    m, n = d_matrix_local.shape
    acc = torch.empty((4, n), dtype=torch.float64, device=d_matrix_local.device)
    prm = params if params is not None else score_params()
    with torch.cuda.device(d_matrix_local.device):
        st = lib.rocco_b200_score_partial_dev(
            ctypes.c_void_p(d_matrix_local.data_ptr()), 1 if d_matrix_local.dtype == torch.float32 else 0, m, n,
            ctypes.byref(prm), ctypes.c_void_p(acc.data_ptr()), ctypes.c_void_p(_stream_ptr(d_matrix_local.device)))
    if st == _lib.ST_NONFINITE:
        raise ValueError("Locus scoring produced non-finite values (or `chrom_matrix` contains non-finite values)")
    _lib.check(st, "score_partial")
    return acc


def score_finalize_device(acc, m_total: int, params=None, details: bool = False):
    torch = _torch()
    lib = _lib.load()
    n = acc.shape[1]
    out = _lib.ScoreOutputs()
    res = {"scores": torch.empty(n, dtype=torch.float64, device=acc.device)}
    if details:
        for k in ("mean", "raw_variance", "prior_variance", "moderated_variance", "standard_error"):
            res[k] = torch.empty(n, dtype=torch.float64, device=acc.device)
    for k, v in res.items():
        setattr(out, k, v.data_ptr())
    prm = params if params is not None else score_params()
    with torch.cuda.device(acc.device):
        st = lib.rocco_b200_score_finalize_dev(ctypes.c_void_p(acc.data_ptr()), int(m_total), n, ctypes.byref(prm),
                                               ctypes.byref(out), ctypes.c_void_p(_stream_ptr(acc.device)))
    if st == _lib.ST_NONFINITE:
        raise ValueError("Locus scoring produced non-finite values")
    _lib.check(st, "score_finalize")
    return res if details else res["scores"]


def score_loci_wls_sample_sharded(d_matrix_local, m_total: int, params=None, group=None, details: bool = False):
    """score_loci_wls with the SAMPLES sharded over ranks (every rank holds [m_local, bins] of the same bins):
    per-sample stages stay local; one all-reduce(sum, float64) of the [4, bins] accumulators; every rank then
    finishes the per-bin combine."""
    import torch.distributed as dist
    acc = score_partial_device(d_matrix_local, params)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(acc, group=group)
    return score_finalize_device(acc, m_total, params, details=details)


# ------------------------------------------------------------------ budget null on device tensors
def budget_null_device(d_centered, d_observed_scores=None, params=None, dependence_lag_hint=None, num_null_draws=25,
                       random_seed=0, min_null_draws=None, stability_abs_tol=5.0e-3, stability_rel_tol=5.0e-2):
    """``estimate_budget_nonnull_fraction_from_wild_bootstrap_null`` (inference.py:988-1148) on CUDA tensors: the centred
    matrix [samples, bins] (float64, as ``score_loci_wls_device(..., details=True)`` returns it) and optionally the observed
    scores stay on the device; returns ``(nonnull_fraction, details)`` with the reference's keys plus the positive-score
    median/count that ``rocco._resolve_chrom_gamma`` needs."""
    from .inference import _budget_details, _hint
    torch = _torch()
    lib = _lib.load()
    if d_centered.dim() != 2 or not d_centered.is_cuda or not d_centered.is_contiguous() or d_centered.dtype != torch.float64:
        raise ValueError("`d_centered` must be a contiguous two-dimensional float64 CUDA tensor")
    m, n = d_centered.shape
    if m == 0 or n == 0:
        raise ValueError("`centered_matrix` must contain at least one locus")
    if d_observed_scores is not None and (d_observed_scores.numel() != n or d_observed_scores.dtype != torch.float64
                                          or not d_observed_scores.is_cuda or not d_observed_scores.is_contiguous()):
        raise ValueError("`observed_scores` must have the same number of loci as `centered_matrix`")
    prm = _lib.BudgetParams()
    lib.rocco_b200_default_budget_params(ctypes.byref(prm))
    if params is not None:
        prm.score = params
    prm.dependence_lag_hint = _hint(dependence_lag_hint)
    prm.num_null_draws = int(max(1, num_null_draws))
    prm.min_null_draws = 0 if min_null_draws is None else int(max(1, min_null_draws))
    prm.stability_abs_tol = float(stability_abs_tol)
    prm.stability_rel_tol = float(stability_rel_tol)
    prm.random_seed = int(random_seed) & 0xFFFFFFFFFFFFFFFF
    res = _lib.BudgetResult()
    dev = d_centered.device
    with torch.cuda.device(dev):
        st = lib.rocco_b200_budget_nonnull_fraction_dev(
            ctypes.c_void_p(d_centered.data_ptr()), m, n,
            None if d_observed_scores is None else ctypes.c_void_p(d_observed_scores.data_ptr()),
            ctypes.byref(prm), ctypes.byref(res), ctypes.c_void_p(_stream_ptr(dev)))
    if st == _lib.ST_NONFINITE:
        raise ValueError("Budget initialization produced non-finite values")
    _lib.check(st, "budget null")
    return float(res.nonnull_fraction), _budget_details(res)

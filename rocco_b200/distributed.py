"""Multi-GPU plumbing: one process per GPU, chromosomes sharded across ranks.

Every chromosome is scored, budgeted and solved independently in the reference (rocco.py:948-1098,
1157-1182), so the data path has NO collective.  The only exchange is a single all-reduce of
[selected bins, bins] for the genome-wide selected fraction (reporting) and a gather of the per-rank
BED shards to rank 0, which then sorts/merges them exactly like combine_chrom_results.
Works with NCCL (GPUs) and gloo (CPU tests) alike; with world size 1 nothing is initialised.
"""
from __future__ import annotations

from typing import Sequence

from .pipeline import lpt_partition
from .rocco import _merge_bed_records


def _dist():
    import torch.distributed as dist
    return dist


def world() -> tuple[int, int]:
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_chromosomes(bin_counts: Sequence[int]) -> list[int]:
    """Indices of the chromosomes this rank owns (LPT packing by bin count, identical on every rank)."""
    rank, size = world()
    return lpt_partition(bin_counts, size)[rank]


def allreduce_selected(selected: int, bins: int, device=None) -> tuple[int, int]:
    """Genome-wide (selected bins, bins): the path's single cross-rank exchange."""
    import torch
    dist = _dist()
    rank, size = world()
    if size == 1:
        return int(selected), int(bins)
    t = torch.tensor([int(selected), int(bins)], dtype=torch.int64, device=device or "cpu")
    dist.all_reduce(t)
    return int(t[0].item()), int(t[1].item())


def gather_bed_records(records: list[tuple[str, int, int]]):
    """Rank 0 receives every rank's records and returns them merged in combine_chrom_results order
    (chromosome string, start, end); other ranks return None."""
    dist = _dist()
    rank, size = world()
    if size == 1:
        return _merge_bed_records(records)
    gathered = [None] * size if rank == 0 else None
    dist.gather_object(records, gathered, dst=0)
    if rank != 0:
        return None
    flat = [r for part in gathered for r in part]
    return _merge_bed_records(flat)


def gather_runs(chrom_index, starts, ends, device=None):
    """All ranks' runs (global chromosome index, first bin, last bin + 1) on rank 0, as three int64 arrays ordered by
    (chromosome index, start); other ranks get None.  Tensor collectives only (one all-gather of the counts, one of the
    padded records), so it runs on NCCL with CUDA tensors and on gloo with CPU tensors alike."""
    import numpy as np
    import torch
    dist = _dist()
    rank, size = world()
    rec = np.stack([np.asarray(chrom_index, dtype=np.int64), np.asarray(starts, dtype=np.int64),
                    np.asarray(ends, dtype=np.int64)]) if len(starts) else np.zeros((3, 0), dtype=np.int64)
    if size > 1:
        dev = device or "cpu"
        cnt = torch.tensor([rec.shape[1]], dtype=torch.int64, device=dev)
        counts = [torch.zeros_like(cnt) for _ in range(size)]
        dist.all_gather(counts, cnt)
        counts = [int(c.item()) for c in counts]
        kmax = max(max(counts), 1)
        pad = torch.zeros((3, kmax), dtype=torch.int64, device=dev)
        pad[:, :rec.shape[1]] = torch.from_numpy(rec).to(dev)
        parts = [torch.zeros_like(pad) for _ in range(size)]
        dist.all_gather(parts, pad)
        if rank != 0:
            return None
        rec = np.concatenate([p[:, :k].cpu().numpy() for p, k in zip(parts, counts)], axis=1)
    order = np.lexsort((rec[1], rec[0]))
    return rec[0][order], rec[1][order], rec[2][order]

"""Synthetic count matrices of the named benchmark shapes.

Per chromosome c: seed = base_seed*1000 + index(c).  Background rate ~ Gamma(0.6, 1.0);
ceil(n/500) planted peaks (width U{4..29} bins, added rate U(5, 40)); per-sample depth
U(0.5, 1.5); counts ~ Poisson(rate * depth); per-sample normalisation U(0.2, 0.5); rounded
to 5 decimals (mimics the reference's readtracks.py:495,517); C-contiguous [samples, bins].

`chrom_matrix_numpy` is the seeded CPU generator used when the same bytes must be fed to the
oracle; `chrom_matrix_torch` draws the same distribution on the device for throughput runs.
"""
from __future__ import annotations

import math

import numpy as np

# rocco/hg38.sizes:1-24 and rocco/hg_params.csv (chrom, budget, gamma) of the reference, as data
HG38_SIZES = {
    "chr1": 248956422, "chr2": 242193529, "chr3": 198295559, "chr4": 190214555, "chr5": 181538259,
    "chr6": 170805979, "chr7": 159345973, "chr8": 145138636, "chr9": 138394717, "chr10": 133797422,
    "chr11": 135086622, "chr12": 133275309, "chr13": 114364328, "chr14": 107043718, "chr15": 101991189,
    "chr16": 90338345, "chr17": 83257441, "chr18": 80373285, "chr19": 58617616, "chr20": 64444167,
    "chr21": 46709983, "chr22": 50818468, "chrX": 156040895, "chrY": 57227415,
}
HG_PARAMS = {
    "chr1": (0.03, 1.0), "chr2": (0.02, 1.0), "chr3": (0.02, 1.0), "chr4": (0.02, 1.0), "chr5": (0.02, 1.0),
    "chr6": (0.02, 1.0), "chr7": (0.025, 1.0), "chr8": (0.025, 1.0), "chr9": (0.025, 1.0), "chr10": (0.02, 1.0),
    "chr11": (0.035, 1.0), "chr12": (0.035, 1.0), "chr13": (0.02, 1.0), "chr14": (0.025, 1.0),
    "chr15": (0.03, 1.0), "chr16": (0.03, 1.0), "chr17": (0.04, 1.0), "chr18": (0.02, 1.0),
    "chr19": (0.045, 1.0), "chr20": (0.03, 1.0), "chr21": (0.02, 1.0), "chr22": (0.03, 1.0),
    "chrX": (0.015, 1.0), "chrY": (0.001, 1.0),
}


def chrom_bins(chrom: str, step: int = 50) -> int:
    return int(math.ceil(HG38_SIZES[chrom] / step))


def chrom_seed(chrom: str, base_seed: int = 0) -> int:
    return base_seed * 1000 + list(HG38_SIZES).index(chrom)


def chrom_matrix_numpy(n_samples: int, n_bins: int, seed: int = 0, dtype=np.float64) -> np.ndarray:
    rng = np.random.default_rng(seed)
    rate = rng.gamma(0.6, 1.0, size=n_bins)
    n_peaks = int(math.ceil(n_bins / 500))
    starts = rng.integers(0, max(n_bins, 1), size=n_peaks)
    widths = rng.integers(4, 30, size=n_peaks)
    heights = rng.uniform(5.0, 40.0, size=n_peaks)
    for s, w, h in zip(starts, widths, heights):
        rate[s:s + w] += h
    depth = rng.uniform(0.5, 1.5, size=(n_samples, 1))
    counts = rng.poisson(rate[None, :] * depth).astype(np.float64)
    norm = rng.uniform(0.2, 0.5, size=(n_samples, 1))
    return np.ascontiguousarray(np.round(counts * norm, 5), dtype=dtype)


def chrom_matrix_torch(n_samples: int, n_bins: int, seed: int, device, dtype=None, sample_stream: int = 0):
    """Same distribution drawn with torch's device RNG (not bit-identical to the NumPy generator).

    ``sample_stream`` != 0 keeps the chromosome's background rate and peaks (functions of ``seed``) but draws the samples
    from another stream: ranks that shard the SAMPLES of one chromosome pass their rank + 1."""
    import torch

    dtype = dtype or torch.float64
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    conc = torch.full((n_bins,), 0.6, device=device, dtype=torch.float32)
    rate = torch._standard_gamma(conc, generator=g)
    # planted peaks: drawn and accumulated on the host (float atomics on the device would make the rates, and
    # through them the Poisson draws, differ from run to run), then prefix-summed
    n_peaks = int(math.ceil(n_bins / 500))
    prng = np.random.default_rng(int(seed) + 7919)
    starts = prng.integers(0, max(n_bins, 1), size=n_peaks)
    widths = prng.integers(4, 30, size=n_peaks)
    heights = prng.uniform(5.0, 40.0, size=n_peaks)
    delta_h = np.zeros(n_bins + 32, dtype=np.float64)
    np.add.at(delta_h, starts, heights)
    np.add.at(delta_h, starts + widths, -heights)
    peaks = np.clip(np.cumsum(delta_h)[:n_bins], 0.0, None).astype(np.float32)
    rate = rate + torch.from_numpy(peaks).to(device)
    if sample_stream:
        g.manual_seed(int(seed) * 1000003 + int(sample_stream))
    depth = torch.empty((n_samples, 1), device=device).uniform_(0.5, 1.5, generator=g)
    norm = torch.empty((n_samples, 1), device=device).uniform_(0.2, 0.5, generator=g)
    out = torch.empty((n_samples, n_bins), device=device, dtype=dtype)
    rows = max(1, (1 << 26) // max(n_bins, 1))
    for r0 in range(0, n_samples, rows):
        r1 = min(n_samples, r0 + rows)
        c = torch.poisson(rate[None, :] * depth[r0:r1], generator=g)
        out[r0:r1] = torch.round(c * norm[r0:r1] * 1e5) / 1e5
    return out

// Mask -> merged intervals on the device (run-boundary detection + ordered compaction).
//
// Replaces the per-bin Python loop of /root/reference/rocco/rocco.py:179-191 and the sort/merge of
// rocco.py:74-95 for the records of one chromosome: bins with solution > 0.5 are emitted as
// (intervals[i], intervals[i+1]) for i in range(len - 1) -- the LAST bin is never emitted -- and
// records with start <= previous end are merged, i.e. every maximal run of selected bins among
// bins 0..n-2 becomes one interval; runs shorter than min_length_bp are dropped afterwards.
//
// Several chromosomes are handled in one pass over their concatenated masks: because the last bin
// of each chromosome is dropped, runs can never join across a chromosome boundary.
#include "common.cuh"

#include <algorithm>

namespace rb {
namespace bed {

constexpr int THREADS = 256;
constexpr int ITEMS = 16;
constexpr int TILE = THREADS * ITEMS;

struct Seg { long long offset; long long n; };

template <bool WRITE>
__global__ void __launch_bounds__(THREADS) k_runs(const uint8_t *mask, long long total, const long long *last_bins,
                                                  int nlast, int2 *tile_counts, const long long *tile_off_start,
                                                  const long long *tile_off_end, long long *starts, long long *ends)
{
    __shared__ uint8_t s_m[TILE + 2];
    __shared__ int s_ws[THREADS / 32], s_we[THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const long long t0 = (long long)blockIdx.x * TILE;
    for (int e = tid; e < TILE + 2; e += THREADS) {
        const long long g = t0 - 1 + e;
        s_m[e] = (g >= 0 && g < total) ? (mask[g] != 0) : 0;
    }
    __syncthreads();
    // dropped (last-of-chromosome) positions that touch [t0-1, t0+TILE] are cleared in place: any number of chromosomes
    // or contigs per tile (thousands of scaffolds of a few bins each included)
    for (int k = tid; k < nlast; k += THREADS) {
        const long long e = last_bins[k] - (t0 - 1);
        if (e >= 0 && e < TILE + 2) s_m[e] = 0;
    }
    __syncthreads();
    unsigned sb = 0, eb = 0;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const int e = tid * ITEMS + j + 1;
        const int cur = s_m[e];
        if (cur && !s_m[e - 1]) sb |= 1u << j;
        if (cur && !s_m[e + 1]) eb |= 1u << j;
    }
    int ns = __popc(sb), ne = __popc(eb);
    int ps = ns, pe = ne;                                   // inclusive warp scans
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int a = __shfl_up_sync(0xffffffffu, ps, d), b = __shfl_up_sync(0xffffffffu, pe, d);
        if (lane >= d) { ps += a; pe += b; }
    }
    if (lane == 31) { s_ws[wid] = ps; s_we[wid] = pe; }
    __syncthreads();
    int bs = 0, be = 0, tots = 0, tote = 0;
    for (int w = 0; w < THREADS / 32; ++w) {
        if (w < wid) { bs += s_ws[w]; be += s_we[w]; }
        tots += s_ws[w]; tote += s_we[w];
    }
    if (!WRITE) {
        if (tid == 0) tile_counts[blockIdx.x] = make_int2(tots, tote);
        return;
    }
    long long os = tile_off_start[blockIdx.x] + bs + (ps - ns);
    long long oe = tile_off_end[blockIdx.x] + be + (pe - ne);
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const long long g = t0 + tid * ITEMS + j;
        if ((sb >> j) & 1u) starts[os++] = g;
        if ((eb >> j) & 1u) ends[oe++] = g + 1;            // half-open
    }
}

__global__ void __launch_bounds__(1024) k_scan_tiles(const int2 *tile_counts, int ntiles, long long *off_s,
                                                      long long *off_e, long long *totals)
{
    // single block: chunked exclusive scan, fixed order
    __shared__ long long s_a[1024], s_b[1024];
    __shared__ long long s_ca, s_cb;
    const int tid = threadIdx.x;
    if (tid == 0) { s_ca = 0; s_cb = 0; }
    __syncthreads();
    for (int base = 0; base < ntiles; base += 1024) {
        const int t = base + tid;
        long long a = 0, b = 0;
        if (t < ntiles) { a = tile_counts[t].x; b = tile_counts[t].y; }
        s_a[tid] = a; s_b[tid] = b;
        __syncthreads();
        for (int d = 1; d < 1024; d <<= 1) {
            long long xa = 0, xb = 0;
            if (tid >= d) { xa = s_a[tid - d]; xb = s_b[tid - d]; }
            __syncthreads();
            s_a[tid] += xa; s_b[tid] += xb;
            __syncthreads();
        }
        if (t < ntiles) { off_s[t] = s_ca + s_a[tid] - a; off_e[t] = s_cb + s_b[tid] - b; }
        __syncthreads();
        if (tid == 1023) { s_ca += s_a[1023]; s_cb += s_b[1023]; }
        __syncthreads();
    }
    if (tid == 0) { totals[0] = s_ca; totals[1] = s_cb; }
}

// Runs over the concatenated masks, as global bin positions [start, end).  Host vectors out.
int runs_batch(const uint8_t *d_mask, const Seg *segs, int nseg, std::vector<long long> &starts,
               std::vector<long long> &ends, cudaStream_t st)
{
    starts.clear(); ends.clear();
    if (nseg <= 0) return 0;
    RB_TRY(ensure_device());
    long long total = 0;
    std::vector<long long> last(nseg);
    for (int k = 0; k < nseg; ++k) {
        if (segs[k].n <= 0) return ST_INVALID;
        last[k] = segs[k].offset + segs[k].n - 1;
        total = std::max(total, segs[k].offset + segs[k].n);
    }
    const int ntiles = (int)((total + TILE - 1) / TILE);
    Arena ar(st);
    long long *d_last = nullptr, *d_os = nullptr, *d_oe = nullptr, *d_tot = nullptr, *d_s = nullptr, *d_e = nullptr;
    int2 *d_cnt = nullptr;
    RB_TRY(ar.alloc(&d_last, nseg));
    RB_TRY(ar.alloc(&d_cnt, ntiles));
    RB_TRY(ar.alloc(&d_os, ntiles));
    RB_TRY(ar.alloc(&d_oe, ntiles));
    RB_TRY(ar.alloc(&d_tot, 2));
    RB_CUDA(cudaMemcpyAsync(d_last, last.data(), sizeof(long long) * nseg, cudaMemcpyHostToDevice, st));
    RB_PROF("mask_to_runs", st, 2.0 * (double)total);
    k_runs<false><<<ntiles, THREADS, 0, st>>>(d_mask, total, d_last, nseg, d_cnt, nullptr, nullptr, nullptr, nullptr);
    RB_LAUNCH_CHECK();
    k_scan_tiles<<<1, 1024, 0, st>>>(d_cnt, ntiles, d_os, d_oe, d_tot);
    RB_LAUNCH_CHECK();
    long long tot[2] = {0, 0};
    RB_CUDA(cudaMemcpyAsync(tot, d_tot, sizeof(tot), cudaMemcpyDeviceToHost, st));
    RB_CUDA(cudaStreamSynchronize(st));
    if (tot[0] != tot[1]) { set_error("run starts (%lld) != run ends (%lld)", tot[0], tot[1]); return ST_CUDA; }
    if (tot[0] == 0) return 0;
    RB_TRY(ar.alloc(&d_s, (size_t)tot[0]));
    RB_TRY(ar.alloc(&d_e, (size_t)tot[0]));
    k_runs<true><<<ntiles, THREADS, 0, st>>>(d_mask, total, d_last, nseg, d_cnt, d_os, d_oe, d_s, d_e);
    RB_LAUNCH_CHECK();
    starts.resize((size_t)tot[0]); ends.resize((size_t)tot[0]);
    RB_CUDA(cudaMemcpyAsync(starts.data(), d_s, sizeof(long long) * tot[0], cudaMemcpyDeviceToHost, st));
    RB_CUDA(cudaMemcpyAsync(ends.data(), d_e, sizeof(long long) * tot[0], cudaMemcpyDeviceToHost, st));
    RB_CUDA(cudaStreamSynchronize(st));
    return 0;
}

}  // namespace bed
}  // namespace rb

using namespace rb;

extern "C" __attribute__((visibility("default"))) long long rocco_b200_mask_to_intervals_dev(
    const uint8_t *d_mask, size_t n, long long first_start, long long step, long long min_length_bp,
    long long *starts_out, long long *ends_out, size_t capacity, void *cuda_stream)
{
    if (!d_mask || n == 0) return ST_INVALID;
    bed::Seg seg{0, (long long)n};
    std::vector<long long> s, e;
    int st = bed::runs_batch(d_mask, &seg, 1, s, e, (cudaStream_t)cuda_stream);
    if (st != 0) return st;
    size_t k = 0;
    for (size_t i = 0; i < s.size(); ++i) {
        const long long a = first_start + s[i] * step, b = first_start + e[i] * step;
        if (min_length_bp > 0 && (b - a) < min_length_bp) continue;
        if (k >= capacity || !starts_out || !ends_out) return ST_INVALID;
        starts_out[k] = a; ends_out[k] = b; ++k;
    }
    return (long long)k;
}

/* Batched form used by the genome pipeline: chromosomes laid out at `offsets` in one mask buffer.
 * Outputs are bin positions relative to each chromosome; chrom_index_out says which chromosome. */
extern "C" __attribute__((visibility("default"))) long long rocco_b200_mask_to_runs_batch_dev(
    const uint8_t *d_mask, const size_t *offsets, const size_t *lengths, int chrom_count,
    long long *start_bin_out, long long *end_bin_out, int *chrom_index_out, size_t capacity, void *cuda_stream)
{
    if (!d_mask || !offsets || !lengths || chrom_count <= 0) return ST_INVALID;
    std::vector<bed::Seg> segs(chrom_count);
    for (int c = 0; c < chrom_count; ++c) segs[c] = bed::Seg{(long long)offsets[c], (long long)lengths[c]};
    for (int c = 1; c < chrom_count; ++c)
        if (segs[c].offset < segs[c - 1].offset + segs[c - 1].n) return ST_INVALID;   // ascending, disjoint
    std::vector<long long> s, e;
    int st = bed::runs_batch(d_mask, segs.data(), chrom_count, s, e, (cudaStream_t)cuda_stream);
    if (st != 0) return st;
    if (s.size() > capacity) return ST_INVALID;
    int c = 0;
    for (size_t i = 0; i < s.size(); ++i) {
        while (c + 1 < chrom_count && s[i] >= segs[c + 1].offset) ++c;
        start_bin_out[i] = s[i] - segs[c].offset;
        end_bin_out[i] = e[i] - segs[c].offset;
        chrom_index_out[i] = c;
    }
    return (long long)s.size();
}

extern "C" __attribute__((visibility("default"))) long long rocco_mask_to_intervals_u8(
    const uint8_t *mask, size_t n, long long first_start, long long step, long long min_length_bp,
    long long *starts_out, long long *ends_out, size_t capacity)
{
    if (!mask || n == 0) return ST_INVALID;
    RB_TRY(ensure_device());
    HostScope lease(true);
    cudaStream_t st = lease.stream();
    Arena ar(st);
    uint8_t *d_m = nullptr;
    RB_TRY(ar.alloc(&d_m, n));
    const void *src = mask;
    if (void *stage = lease.staging(n)) {
        memcpy(stage, mask, n);
        RB_TRY(pull_from_pinned(d_m, stage, n, st));
    } else {
        RB_CUDA(cudaMemcpyAsync(d_m, src, n, cudaMemcpyHostToDevice, st));
    }
    return rocco_b200_mask_to_intervals_dev(d_m, n, first_start, step, min_length_bp, starts_out, ends_out, capacity, st);
}

// ------------------------------------------------------------------ narrowPeak summit offsets
// rocco.py:840-872: for every peak [start, end) the bins whose start lies inside it are searched for the largest WLS mean
// (NaN ignored, first occurrence wins, at least one finite value required); the summit is that bin's centre, reported as
// an offset clipped to the peak.  One warp per peak.
namespace rb {
namespace bed {

__device__ __forceinline__ long long lower_bound_ll(const long long *a, long long n, long long key)
{
    long long lo = 0, hi = n;
    while (lo < hi) { const long long mid = (lo + hi) >> 1; if (a[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;
}

__global__ void __launch_bounds__(256) k_summit_offsets(const long long *__restrict__ starts, const long long *__restrict__ centers,
                                                        const float *__restrict__ mean, long long n_track,
                                                        const long long *__restrict__ pstart, const long long *__restrict__ pend,
                                                        long long n_peaks, long long *__restrict__ out)
{
    const long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (p >= n_peaks) return;
    const long long s = pstart[p], e = pend[p], len = e - s;
    long long result = -1;
    if (len > 0 && n_track > 0) {
        const long long left = lower_bound_ll(starts, n_track, s), right = lower_bound_ll(starts, n_track, e);
        double best = -INFINITY;
        long long best_i = -1;
        bool finite = false;
        for (long long i = left + lane; i < right; i += 32) {
            const double v = (double)mean[i];
            if (isfinite(v)) finite = true;
            if (!isnan(v) && (best_i < 0 || v > best)) { best = v; best_i = i; }       // ascending i: first occurrence kept
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const double ob = __shfl_down_sync(0xffffffffu, best, d);
            const long long oi = __shfl_down_sync(0xffffffffu, best_i, d);
            if (oi >= 0 && (best_i < 0 || ob > best || (ob == best && oi < best_i))) { best = ob; best_i = oi; }
        }
        const unsigned any = __ballot_sync(0xffffffffu, finite);
        if (lane == 0 && any && best_i >= 0) {
            const long long off = centers[best_i] - s, cap = len - 1 > 0 ? len - 1 : 0;
            result = off < 0 ? 0 : (off > cap ? cap : off);
        }
    }
    if (lane == 0) out[p] = result;
}

}  // namespace bed
}  // namespace rb

extern "C" __attribute__((visibility("default"))) int rocco_b200_summit_offsets_dev(
    const long long *d_track_starts, const long long *d_track_centers, const float *d_track_mean, size_t n_track,
    const long long *d_peak_starts, const long long *d_peak_ends, size_t n_peaks, long long *d_offsets_out, void *cuda_stream)
{
    if (n_peaks == 0) return 0;
    if (!d_peak_starts || !d_peak_ends || !d_offsets_out || (n_track && (!d_track_starts || !d_track_centers || !d_track_mean)))
        return ST_INVALID;
    const unsigned grid = (unsigned)((n_peaks * 32 + 255) / 256);
    rb::bed::k_summit_offsets<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(d_track_starts, d_track_centers, d_track_mean,
                                                                          (long long)n_track, d_peak_starts, d_peak_ends,
                                                                          (long long)n_peaks, d_offsets_out);
    RB_LAUNCH_CHECK();
    return 0;
}

extern "C" __attribute__((visibility("default"))) int rocco_narrowpeak_summit_offsets_f32(
    const long long *track_starts, const long long *track_centers, const float *track_mean, size_t n_track,
    const long long *peak_starts, const long long *peak_ends, size_t n_peaks, long long *offsets_out)
{
    if (n_peaks == 0) return 0;
    if (!peak_starts || !peak_ends || !offsets_out) return ST_INVALID;
    RB_TRY(ensure_device());
    HostScope lease(true);
    cudaStream_t st = lease.stream();
    Arena ar(st);
    long long *d_s = nullptr, *d_c = nullptr, *d_ps = nullptr, *d_pe = nullptr, *d_o = nullptr;
    float *d_m = nullptr;
    RB_TRY(ar.alloc(&d_s, n_track)); RB_TRY(ar.alloc(&d_c, n_track)); RB_TRY(ar.alloc(&d_m, n_track));
    RB_TRY(ar.alloc(&d_ps, n_peaks)); RB_TRY(ar.alloc(&d_pe, n_peaks)); RB_TRY(ar.alloc(&d_o, n_peaks));
    if (n_track) {
        RB_CUDA(cudaMemcpyAsync(d_s, track_starts, n_track * sizeof(long long), cudaMemcpyHostToDevice, st));
        RB_CUDA(cudaMemcpyAsync(d_c, track_centers, n_track * sizeof(long long), cudaMemcpyHostToDevice, st));
        RB_CUDA(cudaMemcpyAsync(d_m, track_mean, n_track * sizeof(float), cudaMemcpyHostToDevice, st));
    }
    RB_CUDA(cudaMemcpyAsync(d_ps, peak_starts, n_peaks * sizeof(long long), cudaMemcpyHostToDevice, st));
    RB_CUDA(cudaMemcpyAsync(d_pe, peak_ends, n_peaks * sizeof(long long), cudaMemcpyHostToDevice, st));
    RB_TRY(rocco_b200_summit_offsets_dev(d_s, d_c, d_m, n_track, d_ps, d_pe, n_peaks, d_o, st));
    RB_CUDA(cudaMemcpyAsync(offsets_out, d_o, n_peaks * sizeof(long long), cudaMemcpyDeviceToHost, st));
    RB_CUDA(cudaStreamSynchronize(st));
    return 0;
}

// Matrix assembly / input staging on the device.
//
// Replaces, for decoded alignment records,
//   /root/reference/rocco/native/ccounts_backend.c:2416-2574   per-read coverage loop (delta buffer + prefix sum)
//   /root/reference/rocco/readtracks.py:492-518                scale, trim to the positive range, np.round
//   /root/reference/rocco/readtracks.py:590-633                union of the samples' interval grids, scatter into [m, n]
// The reference builds the float64 matrix on the host and the hot path would then upload 8 bytes per sample-bin; here the
// records (or the float32 coverage) are what crosses PCIe and the matrix is born in HBM in the layout scoring reads.
#include "common.cuh"

#include <algorithm>

namespace rb {
namespace assemble {

constexpr int BAM_FPROPER_PAIR = 2, BAM_FMUNMAP = 8, BAM_FREVERSE = 16, BAM_FREAD2 = 128;

// One thread per record: the reference's filters and fragment arithmetic verbatim in meaning (ccounts_backend.c:2416-2540);
// +1 / -1 on an int32 delta buffer (the reference adds +-1.0f to a float buffer: the same integers below 2^24), or +1 on the
// midpoint bin in one-read-per-bin mode.
__global__ void __launch_bounds__(256) k_count_reads(const int64_t *__restrict__ pos, const int64_t *__restrict__ endpos,
                                                     const uint16_t *__restrict__ flag, const uint8_t *__restrict__ mapq,
                                                     const int64_t *__restrict__ isize, const uint8_t *__restrict__ same_tid,
                                                     size_t n_reads, rocco_b200_count_options opt, long long start, long long end,
                                                     long long step, int *delta, int *direct, unsigned long long count_len)
{
    const long long min_tlen = opt.min_template_length >= 0 ? opt.min_template_length : opt.read_length;
    for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += (size_t)gridDim.x * blockDim.x) {
        const int f = flag[r];
        if (opt.flag_include > 0 && (f & opt.flag_include) != opt.flag_include) continue;
        if ((f & opt.flag_exclude) != 0) continue;
        if ((int)mapq[r] < opt.min_mapping_quality) continue;
        const long long rs = pos[r], re = endpos[r];
        long long a, b;
        if (opt.paired_end_mode > 0) {
            if ((f & BAM_FPROPER_PAIR) == 0 || (f & BAM_FREAD2) != 0) continue;
            if ((f & BAM_FMUNMAP) != 0 || !same_tid[r]) continue;
            const long long tlen = isize[r];
            const long long atl = tlen >= 0 ? tlen : -tlen;
            if (atl == 0 || atl < min_tlen) continue;
            if (opt.max_insert_size > 0 && atl > opt.max_insert_size) continue;
            if (tlen >= 0) { a = rs; b = rs + atl; } else { b = re; a = b - atl; }
            if ((f & BAM_FREVERSE) == 0) { a += opt.shift_forward_strand53; b += opt.shift_forward_strand53; }
            else { a -= opt.shift_reverse_strand53; b -= opt.shift_reverse_strand53; }
        } else if ((f & BAM_FREVERSE) == 0) {
            const long long five = rs + opt.shift_forward_strand53;
            if (opt.extend_bp > 0) { a = five; b = five + opt.extend_bp; }
            else { a = rs + opt.shift_forward_strand53; b = re + opt.shift_forward_strand53; }
        } else {
            const long long five = (re - 1) - opt.shift_reverse_strand53;
            if (opt.extend_bp > 0) { b = five + 1; a = b - opt.extend_bp; }
            else { a = rs - opt.shift_reverse_strand53; b = re - opt.shift_reverse_strand53; }
        }
        if (b <= start || a >= end) continue;
        if (a < start) a = start;
        if (b > end) b = end;
        if (opt.one_read_per_bin) {
            const unsigned long long idx = (unsigned long long)(((a + b) / 2 - start) / step);
            if (idx < count_len) atomicAdd(direct + idx, 1);
            continue;
        }
        unsigned long long i0 = (unsigned long long)((a - start) / step), i1 = (unsigned long long)(((b - 1) - start) / step);
        if (i0 >= count_len) continue;
        if (i1 >= count_len) i1 = count_len - 1;
        if (i0 > i1) continue;
        atomicAdd(delta + i0, 1);
        atomicAdd(delta + i1 + 1, -1);
    }
}

// Prefix sum of the delta buffer in three small kernels (tile sums, scan of the tile sums, tile scan + offset) and the
// conversion to the float32 coverage the reference returns.
constexpr int SC_THREADS = 256, SC_ITEMS = 16, SC_TILE = SC_THREADS * SC_ITEMS;

__device__ __forceinline__ int block_exclusive_scan(int v, int *s_warp, int &total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = lane < SC_THREADS / 32 ? s_warp[lane] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, w, d); if (lane >= d) w += o; }
        if (lane < SC_THREADS / 32) s_warp[lane] = w;
    }
    __syncthreads();
    total = s_warp[SC_THREADS / 32 - 1];
    const int before = wid ? s_warp[wid - 1] : 0;
    __syncthreads();
    return before + inc - v;
}

__global__ void __launch_bounds__(SC_THREADS) k_tile_sums(const int *__restrict__ delta, unsigned long long n, int *tile_sum)
{
    __shared__ int s_warp[32];
    const unsigned long long base = (unsigned long long)blockIdx.x * SC_TILE + (unsigned long long)threadIdx.x * SC_ITEMS;
    int v = 0;
#pragma unroll
    for (int j = 0; j < SC_ITEMS; ++j) if (base + j < n) v += delta[base + j];
    int total;
    block_exclusive_scan(v, s_warp, total);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SC_THREADS) k_scan_tile_sums(int *tile_sum, int n_tiles)
{
    __shared__ int s_warp[32];
    int carry = 0;
    for (int t0 = 0; t0 < n_tiles; t0 += SC_THREADS) {
        const int t = t0 + threadIdx.x;
        const int v = t < n_tiles ? tile_sum[t] : 0;
        int total;
        const int ex = block_exclusive_scan(v, s_warp, total);
        if (t < n_tiles) tile_sum[t] = carry + ex;
        carry += total;
    }
}

__global__ void __launch_bounds__(SC_THREADS) k_tile_scan(const int *__restrict__ delta, const int *__restrict__ direct,
                                                          const int *__restrict__ tile_offset, unsigned long long n, float *counts)
{
    __shared__ int s_warp[32];
    const unsigned long long base = (unsigned long long)blockIdx.x * SC_TILE + (unsigned long long)threadIdx.x * SC_ITEMS;
    int d[SC_ITEMS], v = 0;
#pragma unroll
    for (int j = 0; j < SC_ITEMS; ++j) { d[j] = (base + j < n) ? delta[base + j] : 0; v += d[j]; }
    int total;
    int run = tile_offset[blockIdx.x] + block_exclusive_scan(v, s_warp, total);
#pragma unroll
    for (int j = 0; j < SC_ITEMS; ++j) {
        run += d[j];
        if (base + j < n) counts[base + j] = (float)(run + direct[base + j]);      // coverage + one-read-per-bin hits (ccounts_backend.c:2560-2566)
    }
}

// readtracks.py:492-503: the scaled value of one bin, operation by operation (float32 count widened to float64)
__device__ __forceinline__ double scaled_value(float count, double norm_scale, int scale_by_step, double step, double const_scale)
{
    double v = __dmul_rn((double)count, norm_scale);
    if (scale_by_step) v = __ddiv_rn(v, step);
    if (const_scale >= 0.0) v = __dmul_rn(v, const_scale);
    return v;
}

__global__ void __launch_bounds__(256) k_positive_range(rocco_b200_track T, double step, long long *range /* [2]: min first, max last */)
{
    long long lo = LLONG_MAX, hi = -1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < T.count_len; i += (size_t)gridDim.x * blockDim.x)
        if (scaled_value(T.d_counts[i], T.norm_scale, T.scale_by_step, step, T.const_scale) > 0.0) { lo = min(lo, (long long)i); hi = max(hi, (long long)i); }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, d));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, d));
    }
    if ((threadIdx.x & 31) == 0) {
        if (lo != LLONG_MAX) atomicMin(range, lo);
        if (hi >= 0) atomicMax(range + 1, hi);
    }
}

struct FillRow { const float *counts; long long count_start, first_bin, last_bin; double norm_scale, const_scale; int scale_by_step; int pad; };

// One thread per column: locate the column's genomic position through the segment table, then write every row's value
// (coalesced along the columns for each row).  np.round(v, d) is rint(v * 10^d) / 10^d (NumPy's own formulation).
template <typename OUT>
__global__ void __launch_bounds__(256) k_fill_matrix(const FillRow *__restrict__ rows, int n_rows, const long long *__restrict__ seg_start,
                                                     const long long *__restrict__ seg_col, int n_seg, unsigned long long n_cols,
                                                     long long step, double pow10, OUT *matrix, long long *intervals)
{
    const unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_cols) return;
    int lo = 0, hi = n_seg;                                   // last segment whose first column is <= j
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if ((unsigned long long)seg_col[mid] <= j) lo = mid; else hi = mid; }
    const long long p = seg_start[lo] + (long long)(j - (unsigned long long)seg_col[lo]) * step;
    intervals[j] = p;
    const double dstep = (double)step;
    for (int r = 0; r < n_rows; ++r) {
        const FillRow R = rows[r];
        const long long b = (p - R.count_start) / step;
        double v = 0.0;
        if (p >= R.count_start && b >= R.first_bin && b <= R.last_bin) {
            v = scaled_value(R.counts[b], R.norm_scale, R.scale_by_step, dstep, R.const_scale);
            v = __ddiv_rn(rint(__dmul_rn(v, pow10)), pow10);
        }
        matrix[(unsigned long long)r * n_cols + j] = (OUT)v;
    }
}

}  // namespace assemble
}  // namespace rb

using namespace rb;
#define RB_API extern "C" __attribute__((visibility("default")))

RB_API int rocco_b200_count_alignment_region_dev(const int64_t *d_pos, const int64_t *d_endpos, const uint16_t *d_flag,
                                                 const uint8_t *d_mapq, const int64_t *d_isize, const uint8_t *d_mate_same_tid,
                                                 size_t n_reads, const rocco_b200_count_options *options, int64_t start, int64_t end,
                                                 int64_t step, float *d_counts, size_t count_len, void *cuda_stream)
{
    if (!options || !d_counts || step <= 0 || end <= start || count_len == 0) return ST_INVALID;
    if (n_reads && (!d_pos || !d_endpos || !d_flag || !d_mapq || !d_isize || !d_mate_same_tid)) return ST_INVALID;
    RB_TRY(ensure_device());
    cudaStream_t st = (cudaStream_t)cuda_stream;
    Arena ar(st);
    int *d_delta = nullptr, *d_direct = nullptr, *d_tiles = nullptr;
    const int n_tiles = (int)((count_len + assemble::SC_TILE - 1) / assemble::SC_TILE);
    RB_TRY(ar.alloc(&d_delta, count_len + 1));
    RB_TRY(ar.alloc(&d_direct, count_len));
    RB_TRY(ar.alloc(&d_tiles, (size_t)n_tiles));
    RB_CUDA(cudaMemsetAsync(d_delta, 0, sizeof(int) * (count_len + 1), st));
    RB_CUDA(cudaMemsetAsync(d_direct, 0, sizeof(int) * count_len, st));
    if (n_reads) {
        const unsigned grid = (unsigned)std::min<size_t>((n_reads + 255) / 256, (size_t)sm_count() * 16);
        RB_PROF("assemble_count_reads", st, (double)n_reads * 28.0);
        assemble::k_count_reads<<<grid, 256, 0, st>>>(d_pos, d_endpos, d_flag, d_mapq, d_isize, d_mate_same_tid, n_reads, *options,
                                                      (long long)start, (long long)end, (long long)step, d_delta, d_direct,
                                                      (unsigned long long)count_len);
        RB_LAUNCH_CHECK();
    }
    assemble::k_tile_sums<<<n_tiles, assemble::SC_THREADS, 0, st>>>(d_delta, (unsigned long long)count_len, d_tiles);
    RB_LAUNCH_CHECK();
    assemble::k_scan_tile_sums<<<1, assemble::SC_THREADS, 0, st>>>(d_tiles, n_tiles);
    RB_LAUNCH_CHECK();
    assemble::k_tile_scan<<<n_tiles, assemble::SC_THREADS, 0, st>>>(d_delta, d_direct, d_tiles, (unsigned long long)count_len, d_counts);
    RB_LAUNCH_CHECK();
    return 0;
}

RB_API int rocco_b200_track_positive_range_dev(const rocco_b200_track *tracks, int n_tracks, int64_t step, int64_t *first_bin_out,
                                               int64_t *last_bin_out, void *cuda_stream)
{
    if (!tracks || n_tracks <= 0 || step <= 0 || !first_bin_out || !last_bin_out) return ST_INVALID;
    RB_TRY(ensure_device());
    cudaStream_t st = (cudaStream_t)cuda_stream;
    Arena ar(st);
    long long *d_range = nullptr;
    RB_TRY(ar.alloc(&d_range, (size_t)2 * n_tracks));
    std::vector<long long> init((size_t)2 * n_tracks);
    for (int t = 0; t < n_tracks; ++t) { init[2 * t] = LLONG_MAX; init[2 * t + 1] = -1; }
    RB_CUDA(cudaMemcpyAsync(d_range, init.data(), sizeof(long long) * init.size(), cudaMemcpyHostToDevice, st));
    for (int t = 0; t < n_tracks; ++t) {
        if (!tracks[t].d_counts || tracks[t].count_len == 0) continue;
        const unsigned grid = (unsigned)std::min<size_t>((tracks[t].count_len + 255) / 256, (size_t)sm_count() * 8);
        assemble::k_positive_range<<<grid, 256, 0, st>>>(tracks[t], (double)step, d_range + 2 * t);
        RB_LAUNCH_CHECK();
    }
    RB_CUDA(cudaMemcpyAsync(init.data(), d_range, sizeof(long long) * init.size(), cudaMemcpyDeviceToHost, st));
    RB_CUDA(cudaStreamSynchronize(st));
    for (int t = 0; t < n_tracks; ++t) {
        const bool none = init[2 * t + 1] < 0;
        first_bin_out[t] = none ? -1 : init[2 * t];
        last_bin_out[t] = none ? -1 : init[2 * t + 1];
    }
    return 0;
}

RB_API int rocco_b200_assemble_matrix_dev(const rocco_b200_track *tracks, const int64_t *first_bin, const int64_t *last_bin,
                                          int n_tracks, int64_t step, int round_digits, const int64_t *segment_start_bp,
                                          const int64_t *segment_first_column, int n_segments, size_t n_columns, int out_f32,
                                          void *d_matrix, int64_t *d_intervals, void *cuda_stream)
{
    if (!tracks || !first_bin || !last_bin || n_tracks <= 0 || step <= 0 || !segment_start_bp || !segment_first_column ||
        n_segments <= 0 || n_columns == 0 || !d_matrix || !d_intervals || round_digits < 0 || round_digits > 15)
        return ST_INVALID;
    RB_TRY(ensure_device());
    cudaStream_t st = (cudaStream_t)cuda_stream;
    std::vector<assemble::FillRow> rows;
    for (int t = 0; t < n_tracks; ++t) {
        if (first_bin[t] < 0 || last_bin[t] < first_bin[t]) continue;             // no data: the track is excluded (readtracks.py:592-598)
        rows.push_back({tracks[t].d_counts, (long long)tracks[t].count_start, (long long)first_bin[t], (long long)last_bin[t],
                        tracks[t].norm_scale, tracks[t].const_scale, tracks[t].scale_by_step, 0});
    }
    if (rows.empty()) return ST_INVALID;
    Arena ar(st);
    assemble::FillRow *d_rows = nullptr;
    long long *d_seg = nullptr;
    RB_TRY(ar.alloc(&d_rows, rows.size()));
    RB_TRY(ar.alloc(&d_seg, (size_t)2 * n_segments));
    RB_CUDA(cudaMemcpyAsync(d_rows, rows.data(), sizeof(assemble::FillRow) * rows.size(), cudaMemcpyHostToDevice, st));
    RB_CUDA(cudaMemcpyAsync(d_seg, segment_start_bp, sizeof(long long) * n_segments, cudaMemcpyHostToDevice, st));
    RB_CUDA(cudaMemcpyAsync(d_seg + n_segments, segment_first_column, sizeof(long long) * n_segments, cudaMemcpyHostToDevice, st));
    double pow10 = 1.0;
    for (int k = 0; k < round_digits; ++k) pow10 *= 10.0;
    const unsigned grid = (unsigned)((n_columns + 255) / 256);
    RB_PROF("assemble_fill_matrix", st, (double)rows.size() * (double)n_columns * (4.0 + (out_f32 ? 4.0 : 8.0)));
    if (out_f32)
        assemble::k_fill_matrix<float><<<grid, 256, 0, st>>>(d_rows, (int)rows.size(), d_seg, d_seg + n_segments, n_segments,
                                                            (unsigned long long)n_columns, (long long)step, pow10, (float *)d_matrix,
                                                            (long long *)d_intervals);
    else
        assemble::k_fill_matrix<double><<<grid, 256, 0, st>>>(d_rows, (int)rows.size(), d_seg, d_seg + n_segments, n_segments,
                                                             (unsigned long long)n_columns, (long long)step, pow10, (double *)d_matrix,
                                                             (long long *)d_intervals);
    RB_LAUNCH_CHECK();
    RB_CUDA(cudaStreamSynchronize(st));            // the host-side row / segment tables go out of scope
    return 0;
}

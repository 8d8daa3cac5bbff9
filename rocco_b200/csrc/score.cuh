// Internal interfaces between the scoring translation units.
#pragma once
#include "common.cuh"

namespace rb {
namespace score {

// Cross-fit Whittaker baseline over `rows` rows of length n (row-major, `row_stride` elements apart).
// log_transform: y = log2(max(x,0)+1) - pilot[row];  write_baseline: out = baseline, else out = y - baseline.
int whittaker_rows(const void *d_x, int in_f32, int log_transform, const double *d_pilot, long long rows, long long n,
                   long long row_stride, double lam, int write_baseline, double *d_out, int *d_bad, cudaStream_t st);

// Pilot offsets: per-row median of log2(max(x,0)+1) -- exact for n <= 4096, else the median of a
// 4096-point strided sample (the offset cancels in  y - baseline(y)  up to the solver's noise).
int pilot_offsets(const void *d_x, int in_f32, long long rows, long long n, long long row_stride, double *d_pilot,
                  cudaStream_t st);

// Centered matrix -> per-locus WLS outputs (wls_backend.c:744-947).
int centered_wls(const double *d_centered, long long m, long long n, const rocco_b200_score_params &prm,
                 rocco_b200_score_outputs *out, cudaStream_t st);

int resolve_spatial_window(long long n, int requested);
int resolve_baseline_window(long long n, int target);
double whittaker_lambda(int block);

}  // namespace score
}  // namespace rb

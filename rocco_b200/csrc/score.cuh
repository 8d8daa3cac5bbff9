// Internal interfaces between the scoring translation units.
#pragma once
#include "common.cuh"

namespace rb {
namespace score {

// Cross-fit Whittaker baseline over `rows` rows of length n (row-major, `row_stride` elements apart).
// log_transform: y = log2(max(x,0)+1) - pilot[row];  write_baseline: out = baseline, else out = y - baseline.
int whittaker_rows(const void *d_x, int in_f32, int log_transform, const double *d_pilot, long long rows, long long n,
                   long long row_stride, double lam, int write_baseline, double *d_out, int *d_bad, cudaStream_t st);

// 0: cluster-pair kernel for the steady tiles (default), 1: single-CTA kernel; returns the previous mode.
int whittaker_set_mode(int mode);

// Pilot offsets: per-row median of log2(max(x,0)+1) -- exact for n <= 4096, else the median of a
// 4096-point strided sample (the offset cancels in  y - baseline(y)  up to the solver's noise).
int pilot_offsets(const void *d_x, int in_f32, long long rows, long long n, long long row_stride, double *d_pilot,
                  cudaStream_t st);

// Centered matrix -> per-locus WLS outputs (wls_backend.c:744-947).
int centered_wls(const double *d_centered, long long m, long long n, const rocco_b200_score_params &prm,
                 rocco_b200_score_outputs *out, cudaStream_t st);

// a / b from a correctly rounded reciprocal rb = RN(1/b): Markstein's final step (q = RN(a*rb), r = a - q*b exactly by
// FMA, RN(q + r*rb)) gives the correctly rounded quotient for operands in the normal range without the generic
// division's special-case slow path (~3 issue slots instead of ~30).
__device__ __forceinline__ double div_rcp(double a, double b, double rb)
{
    const double q = __dmul_rn(a, rb);
    const double r = __fma_rn(-q, b, a);
    return __fma_rn(r, rb, q);
}

// AR(1) innovation variance of one window from its three sums (wls_backend.c:665-712), same association order;
// intrinsics keep nvcc from contracting the products into FMAs.  rwd = RN(1/wd), shrink = 1/(wd+1).
__device__ __forceinline__ double ar1_window_variance(double s1, double s2, double sl, double first, double last, double wd,
                                                      double rwd, double pairs, double shrink)
{
    const double sum_head = __dsub_rn(s1, last);
    const double sum_tail = __dsub_rn(s1, first);
    const double mu = div_rcp(s1, wd, rwd);
    double g0 = __dsub_rn(s2, __dmul_rn(__dmul_rn(wd, mu), mu));
    if (g0 < 0.0) g0 = 0.0;
    double g1 = __dsub_rn(sl, __dmul_rn(mu, sum_head));
    g1 = __dsub_rn(g1, __dmul_rn(mu, sum_tail));
    g1 = __dadd_rn(g1, __dmul_rn(__dmul_rn(pairs, mu), mu));
    const double flo = __dmul_rn(1.0e-4, __dadd_rn(g0, 1.0));
    const double den = __dadd_rn(__dmul_rn(g0, __dadd_rn(1.0, shrink)), flo);
    const double eps = __dmul_rn(1.0e-12, __dadd_rn(g0, 1.0));
    double beta = 0.0;
    if (den > eps) beta = div_rcp(g1, den, __drcp_rn(den));
    if (beta > 0.99) beta = 0.99; else if (beta < 0.0) beta = 0.0;
    const double gam0 = div_rcp(g0, wd, rwd);
    double omb = __dsub_rn(1.0, __dmul_rn(beta, beta));
    if (omb < 0.0) omb = 0.0;
    return fmax(__dmul_rn(gam0, omb), 0.0);
}

int resolve_spatial_window(long long n, int requested);
int resolve_baseline_window(long long n, int target);
double whittaker_lambda(int block);

}  // namespace score
}  // namespace rb

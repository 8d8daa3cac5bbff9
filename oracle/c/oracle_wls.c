/*
 * ORACLE (test infrastructure only -- never linked into or called by the product path).
 *
 * CPU restatement of the reference's EB-moderated WLS locus scoring:
 *   /root/reference/rocco/native/wls_backend.c:232-260  spatial window rule
 *   /root/reference/rocco/native/wls_backend.c:610-742  rolling AR(1) innovation variance
 *   /root/reference/rocco/native/wls_backend.c:394-608  monotone variance-vs-|signal| trend
 *        (lexicographic (x, y) sort 207-230/454, equal-count bins 478-505, PAVA 262-338,
 *         knot de-duplication 547-560, linear interpolation 341-391)
 *   /root/reference/rocco/native/wls_backend.c:744-947  posterior precision + per-locus combine
 *
 * Order statistics are exact, so any correct sort/selection yields the reference's
 * bits; every floating-point expression keeps the reference's association order.
 * The rolling sums are SLIDING updates, as upstream (their drift is part of the
 * reference's answer). Pinned against oracle/_ref by tests/test_oracle_pin.py.
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double x, y; } xy_t;

static int cmp_xy(const void *a, const void *b)
{
    const xy_t *p = (const xy_t *)a, *q = (const xy_t *)b;
    if (p->x != q->x) return (p->x < q->x) ? -1 : 1;
    if (p->y != q->y) return (p->y < q->y) ? -1 : 1;
    return 0;
}

static int cmp_f64(const void *a, const void *b)
{
    const double p = *(const double *)a, q = *(const double *)b;
    return (p < q) ? -1 : (p > q) ? 1 : 0;
}

/* median with the reference's even-count rule 0.5*(lo+hi); sorts `v` in place. wls_backend.c:121-145 */
static double median_inplace(double *v, size_t n)
{
    if (n == 0) return 0.0;
    if (n == 1) return v[0];
    qsort(v, n, sizeof(double), cmp_f64);
    if (n & 1U) return v[n / 2];
    return 0.5 * (v[n / 2 - 1] + v[n / 2]);
}

size_t oracle_resolve_spatial_window(size_t n, int requested)
{
    if (n < 5) return 0;
    size_t w = requested > 0 ? (size_t)requested : 31;
    if (w < 5) w = 5;
    if (w > n) w = n;
    if ((w & 1U) == 0) w = (w == n) ? (w - 1) : (w + 1);
    return (w < 5) ? 0 : w;
}

/* wls_backend.c:176-198 : 1.4826 * MAD floored at 1e-6; destroys `v` */
static double robust_scale(double *v, size_t n)
{
    if (n == 0) return 1.0e-6;
    const double med = median_inplace(v, n);
    for (size_t i = 0; i < n; ++i) v[i] = fabs(v[i] - med);
    double mad = median_inplace(v, n);
    mad *= 1.4826;
    return (mad > 1.0e-6) ? mad : 1.0e-6;
}

/* weighted pool-adjacent-violators, non-decreasing fit. wls_backend.c:262-338 */
static int pava(const double *v, const double *w, size_t n, double *fit)
{
    if (n == 0) return 0;
    double *bv = (double *)malloc(n * sizeof(double));
    double *bw = (double *)malloc(n * sizeof(double));
    size_t *bl = (size_t *)malloc(n * sizeof(size_t));
    if (!bv || !bw || !bl) { free(bv); free(bw); free(bl); return -1; }
    size_t nb = 0;
    for (size_t i = 0; i < n; ++i) {
        bv[nb] = v[i]; bw[nb] = fmax(w[i], 1.0e-8); bl[nb] = 1; ++nb;
        while (nb >= 2 && bv[nb - 2] > bv[nb - 1]) {
            const double tw = bw[nb - 2] + bw[nb - 1];
            const double mv = ((bv[nb - 2] * bw[nb - 2]) + (bv[nb - 1] * bw[nb - 1])) / tw;
            bv[nb - 2] = mv; bw[nb - 2] = tw; bl[nb - 2] += bl[nb - 1];
            --nb;
        }
    }
    size_t k = 0;
    for (size_t b = 0; b < nb; ++b)
        for (size_t r = 0; r < bl[b]; ++r) fit[k++] = bv[b];
    free(bv); free(bw); free(bl);
    return 0;
}

/* wls_backend.c:341-391 */
static double interp_knots(const double *kx, const double *ky, size_t nk, double t)
{
    if (nk == 0) return 1.0e-8;
    if (nk == 1 || t <= kx[0]) return ky[0];
    if (t >= kx[nk - 1]) return ky[nk - 1];
    size_t lo = 0, hi = nk - 1;
    while (hi - lo > 1) {
        const size_t mid = lo + (hi - lo) / 2;
        if (kx[mid] <= t) lo = mid; else hi = mid;
    }
    const double xl = kx[lo], xr = kx[hi];
    if (xr <= xl) return fmax(ky[hi], ky[lo]);
    const double wgt = (t - xl) / (xr - xl);
    return ky[lo] + (wgt * (ky[hi] - ky[lo]));
}

static void fill(double *v, size_t n, double c) { for (size_t i = 0; i < n; ++i) v[i] = c; }

/* wls_backend.c:394-608. Also reports the knots (for stage-by-stage GPU checks). */
int oracle_fit_variance_trend(
    const double *signal, const double *var, size_t n, double *trend,
    double *knot_x_out, double *knot_y_out, int *knot_count_out)
{
    if (knot_count_out) *knot_count_out = 0;
    xy_t *pairs = (xy_t *)malloc((n ? n : 1) * sizeof(xy_t));
    double *scratch = (double *)malloc((n ? n : 1) * sizeof(double));
    if (!pairs || !scratch) { free(pairs); free(scratch); return -1; }

    size_t nf = 0;
    for (size_t i = 0; i < n; ++i) {
        if (isfinite(signal[i]) && isfinite(var[i])) {
            pairs[nf].x = fabs(signal[i]);
            pairs[nf].y = fmax(var[i], 1.0e-8);
            scratch[nf] = pairs[nf].y;
            ++nf;
        }
    }
    double fallback = 1.0e-6;
    if (nf > 0) fallback = fmax(median_inplace(scratch, nf), 1.0e-8);
    if (nf < 4) { fill(trend, n, fallback); free(pairs); free(scratch); return 0; }

    qsort(pairs, nf, sizeof(xy_t), cmp_xy);

    const size_t nbins = (size_t)fmax(4.0, floor(1.0 + (log((double)nf + 1.0) / log(2.0))));
    double *bx = (double *)malloc(6 * nbins * sizeof(double));
    if (!bx) { free(pairs); free(scratch); return -1; }
    double *by = bx + nbins, *bw = by + nbins, *fit = bw + nbins, *kx = fit + nbins, *ky = kx + nbins;

    size_t used = 0;
    for (size_t b = 0; b < nbins; ++b) {
        const size_t lo = (b * nf) / nbins, hi = ((b + 1) * nf) / nbins;
        if (hi <= lo) continue;
        const size_t wd = hi - lo;
        bx[used] = (wd & 1U) ? pairs[lo + wd / 2].x
                             : 0.5 * (pairs[lo + wd / 2 - 1].x + pairs[lo + wd / 2].x);
        for (size_t j = 0; j < wd; ++j) scratch[j] = pairs[lo + j].y;
        by[used] = median_inplace(scratch, wd);
        bw[used] = (double)wd;
        ++used;
    }

    int status = 0;
    if (used == 0) {
        fill(trend, n, fallback);
    } else if (used == 1) {
        fill(trend, n, fmax(by[0], 1.0e-8));
    } else if (pava(by, bw, used, fit) != 0) {
        status = -1;
    } else {
        size_t nk = 0;
        for (size_t b = 0; b < used; ++b) {
            const double cx = bx[b], cy = fmax(fit[b], 1.0e-8);
            if (nk > 0 && cx <= kx[nk - 1]) { ky[nk - 1] = fmax(ky[nk - 1], cy); continue; }
            kx[nk] = cx; ky[nk] = cy; ++nk;
        }
        if (knot_count_out) {
            *knot_count_out = (int)nk;
            if (knot_x_out) memcpy(knot_x_out, kx, nk * sizeof(double));
            if (knot_y_out) memcpy(knot_y_out, ky, nk * sizeof(double));
        }
        if (nk == 0) fill(trend, n, fallback);
        else if (nk == 1) fill(trend, n, fmax(ky[0], 1.0e-8));
        else
            for (size_t i = 0; i < n; ++i)
                trend[i] = isfinite(signal[i])
                               ? fmax(interp_knots(kx, ky, nk, fabs(signal[i])), 1.0e-8)
                               : fallback;
    }
    free(bx); free(pairs); free(scratch);
    return status;
}

/* wls_backend.c:610-742. `w` must already be resolved (odd, >= 5, <= n). */
int oracle_rolling_ar1_variance(const double *v, size_t n, size_t w_req, double *var_out)
{
    if (v == NULL || var_out == NULL || n == 0) return -2;
    const size_t w = oracle_resolve_spatial_window(n, (int)w_req);
    if (w == 0 || n < 4) { memset(var_out, 0, n * sizeof(double)); return 0; }

    const size_t half = w / 2, last = n - w;
    double *at_start = (double *)malloc((last + 1) * sizeof(double));
    if (!at_start) return -1;

    double s1 = 0.0, s2 = 0.0, sl = 0.0;          /* sum y, sum y^2, sum y_k*y_{k+1} */
    for (size_t k = 0; k < w; ++k) {
        const double c = v[k];
        s1 += c;
        s2 += c * c;
        if (k < w - 1) sl += c * v[k + 1];
    }
    const double wd = (double)w, pairs = (double)(w - 1);
    for (size_t t = 0; t <= last; ++t) {
        const double out_v = v[t], in_v = v[t + w - 1];
        const double sum_head = s1 - in_v;        /* all but the last  */
        const double sum_tail = s1 - out_v;       /* all but the first */
        const double mu = s1 / wd;
        double g0 = s2 - (wd * mu * mu);
        const double shrink = 1.0 / (wd + 1.0);
        double beta = 0.0;
        if (g0 < 0.0) g0 = 0.0;
        const double g1 = sl - (mu * sum_head) - (mu * sum_tail) + (pairs * mu * mu);
        const double floor_ = 1.0e-4 * (g0 + 1.0);
        const double den = (g0 * (1.0 + shrink)) + floor_;
        const double eps = 1.0e-12 * (g0 + 1.0);
        if (den > eps) beta = g1 / den;
        if (beta > 0.99) beta = 0.99; else if (beta < 0.0) beta = 0.0;
        const double gam0 = g0 / wd;
        double omb = 1.0 - (beta * beta);
        if (omb < 0.0) omb = 0.0;
        at_start[t] = fmax(gam0 * omb, 0.0);
        if (t < last) {
            const double nx = v[t + w], lag_l = v[t + w - 1], lag_r = v[t + 1];
            s1 = (s1 - out_v) + nx;
            s2 = s2 - (out_v * out_v) + (nx * nx);
            sl = sl - (out_v * lag_r) + (lag_l * nx);
        }
    }
    for (size_t j = 0; j < n; ++j) {
        size_t t = (j < half) ? 0 : (j - half);
        if (t > last) t = last;
        var_out[j] = at_start[t];
    }
    free(at_start);
    return 0;
}

/* wls_backend.c:744-947. Same argument meaning and status codes as rocco_score_centered_wls_f64. */
int oracle_score_centered_wls(
    const double *centered, size_t m, size_t n,
    double lower_bound_z, double prior_df, double min_effect, int use_min_effect,
    int spatial_window, double precision_floor_ratio,
    double *mean_out, double *raw_var_out, double *prior_var_out, double *mod_var_out,
    double *se_out, double *score_out, double *total_df_out, int *window_out)
{
    if (!centered || !mean_out || !raw_var_out || !prior_var_out || !mod_var_out || !se_out || !score_out)
        return -2;
    if (m == 0 || n == 0) return -2;

    const double pdf = fmax(prior_df, 0.0);
    const double pfr = fmax(precision_floor_ratio, 0.0);
    const size_t w = oracle_resolve_spatial_window(n, spatial_window);
    const double ldf = w > 0 ? fmax(4.0, (double)w - 3.0) : 1.0;
    const double tdf = ldf + pdf;
    if (total_df_out) *total_df_out = tdf;
    if (window_out) *window_out = (int)w;

    double *buf = (double *)calloc(7 * n, sizeof(double));
    if (!buf) return -1;
    double *obs = buf, *prior = buf + n, *wsum = buf + 2 * n, *psum = buf + 3 * n;
    double *rsum = buf + 4 * n, *qsum = buf + 5 * n, *tmp = buf + 6 * n;

    for (size_t s = 0; s < m; ++s) {
        const double *row = centered + s * n;
        if (w == 0 || n < 4) {
            memcpy(tmp, row, n * sizeof(double));
            double sc = robust_scale(tmp, n);
            sc = fmax(sc * sc, 1.0e-8);
            fill(obs, n, sc);
            fill(prior, n, sc);
        } else {
            int st = oracle_rolling_ar1_variance(row, n, w, obs);
            if (st != 0) { free(buf); return st == -1 ? -1 : -2; }
            for (size_t j = 0; j < n; ++j) obs[j] = fmax(obs[j], 1.0e-8);
            st = oracle_fit_variance_trend(row, obs, n, prior, NULL, NULL, NULL);
            if (st != 0) { free(buf); return st == -1 ? -1 : -2; }
        }
        for (size_t j = 0; j < n; ++j) {
            const double ov = fmax(obs[j], 1.0e-8), pv = fmax(prior[j], 1.0e-8);
            double post = ((ldf * ov) + (pdf * pv)) / fmax(tdf, 1.0);
            const double flo = pfr * pv;
            if (post < flo) post = flo;
            post = fmax(post, 1.0e-8);
            const double prec = 1.0 / post;
            rsum[j] += 1.0 / ov;
            qsum[j] += 1.0 / pv;
            psum[j] += prec;
            wsum[j] += prec * row[j];
        }
    }
    for (size_t j = 0; j < n; ++j) {
        const double P = fmax(psum[j], 1.0e-8);
        mean_out[j] = wsum[j] / P;
        raw_var_out[j] = (double)m / fmax(rsum[j], 1.0e-8);
        prior_var_out[j] = (double)m / fmax(qsum[j], 1.0e-8);
        mod_var_out[j] = (double)m / P;
        se_out[j] = sqrt(1.0 / P);
        const double z = mean_out[j] / fmax(se_out[j], 1.0e-8);
        score_out[j] = use_min_effect ? (mean_out[j] - fmax(min_effect, 0.0)) / fmax(se_out[j], 1.0e-8)
                                      : z - lower_bound_z;
    }
    free(buf);
    return 0;
}

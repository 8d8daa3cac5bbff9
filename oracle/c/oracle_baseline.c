/*
 * ORACLE (test infrastructure only -- never linked into or called by the product path).
 *
 * CPU restatement of the reference's cross-fit Whittaker baseline:
 *   /root/reference/rocco/native/baseline_backend.c:79-173   pentadiagonal LDL^T solve
 *   /root/reference/rocco/native/baseline_backend.c:175-250  parity-masked system
 *   /root/reference/rocco/native/baseline_backend.c:252-334  even/odd cross-fit average, matrix loop
 * The floating-point operation order of the factorisation and of the three
 * substitution sweeps is preserved so the output is bit-identical to the
 * reference (pinned against oracle/_ref by tests/test_oracle_pin.py).
 *
 * Restated as: (1) a data-independent factor of  W_p + lambda * D'D  that
 * depends only on (n, lambda, parity); (2) forward / diagonal / backward
 * sweeps on the masked right-hand side.
 */
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    double *diag;   /* D of LDL^T,            length n   */
    double *sub1;   /* first sub-diagonal of L, length n-1 */
    double *sub2;   /* second sub-diagonal of L, length n-2 */
} penta_factor;

static double band_main(size_t i, size_t n, int parity, double lam)
{
    const double w = ((i & 1U) == (size_t)parity) ? 1.0 : 0.0;
    if (i == 0 || i == n - 1) return w + lam;
    if (i == 1 || i == n - 2) return w + (5.0 * lam);
    return w + (6.0 * lam);
}

static double band_off1(size_t i, size_t n, double lam)
{
    return (i == 0 || i == n - 2) ? (-2.0 * lam) : (-4.0 * lam);
}

/* n >= 3 here (public entry guarantees n >= 25). baseline_backend.c:106-140 */
static void factor_system(penta_factor *f, size_t n, int parity, double lam)
{
    double *d = f->diag, *l1 = f->sub1, *l2 = f->sub2;
    double t1, t2;

    d[0] = band_main(0, n, parity, lam);
    l1[0] = band_off1(0, n, lam) / d[0];
    l2[0] = lam / d[0];

    d[1] = band_main(1, n, parity, lam) - ((l1[0] * l1[0]) * d[0]);
    t1 = ((l2[0] * d[0]) * l1[0]);
    l1[1] = (band_off1(1, n, lam) - t1) / d[1];
    if (n > 3) l2[1] = lam / d[1];

    for (size_t i = 2; i < n; ++i) {
        t1 = ((l1[i - 1] * l1[i - 1]) * d[i - 1]);
        t2 = ((l2[i - 2] * l2[i - 2]) * d[i - 2]);
        d[i] = band_main(i, n, parity, lam) - t1 - t2;
        if (i + 2 <= n) {
            t1 = ((l2[i - 1] * d[i - 1]) * l1[i - 1]);
            l1[i] = (band_off1(i, n, lam) - t1) / d[i];
        }
        if (i + 3 <= n) l2[i] = lam / d[i];
    }
}

/* Solve in place: x holds the masked rhs on entry, the fit on exit. work: 2n doubles. */
static void solve_factored(const penta_factor *f, size_t n, double *x, double *work)
{
    const double *d = f->diag, *l1 = f->sub1, *l2 = f->sub2;
    double *fw = work, *dz = work + n;
    double t1, t2;

    fw[0] = x[0];
    fw[1] = x[1] - (l1[0] * fw[0]);
    for (size_t i = 2; i < n; ++i) {
        t1 = l1[i - 1] * fw[i - 1];
        t2 = l2[i - 2] * fw[i - 2];
        fw[i] = x[i] - t1 - t2;
    }
    for (size_t i = 0; i < n; ++i) dz[i] = fw[i] / d[i];

    x[n - 1] = dz[n - 1];
    x[n - 2] = dz[n - 2] - (l1[n - 2] * x[n - 1]);
    for (size_t i = n - 2; i-- > 0;) {
        t1 = l1[i] * x[i + 1];
        t2 = l2[i] * x[i + 2];
        x[i] = dz[i] - t1 - t2;
    }
}

static void masked_rhs(const double *y, size_t n, int parity, double *rhs)
{
    /* interior: weight * y (keeps the reference's signed zeros); the two end pairs are selects */
    for (size_t i = 0; i < n; ++i) {
        const int on = ((i & 1U) == (size_t)parity);
        if (i < 2 || i + 2 >= n) rhs[i] = on ? y[i] : 0.0;
        else rhs[i] = (on ? 1.0 : 0.0) * y[i];
    }
}

int oracle_crossfit_whittaker_baseline(const double *y, size_t n, double lam, double *out)
{
    if (y == NULL || out == NULL) return -1;
    if (n < 25) {                       /* baseline_backend.c:266-273 */
        for (size_t i = 0; i < n; ++i) out[i] = 0.0;
        return 0;
    }
    double *buf = (double *)malloc((size_t)6 * n * sizeof(double));
    if (buf == NULL) return -1;
    penta_factor f = { buf, buf + n, buf + 2 * n };
    double *work = buf + 3 * n;         /* 2n */
    double *odd = buf + 5 * n;          /* n  */

    factor_system(&f, n, 0, lam);
    masked_rhs(y, n, 0, out);
    solve_factored(&f, n, out, work);

    factor_system(&f, n, 1, lam);
    masked_rhs(y, n, 1, odd);
    solve_factored(&f, n, odd, work);

    for (size_t i = 0; i < n; ++i) out[i] = 0.5 * (out[i] + odd[i]);
    free(buf);
    return 0;
}

int oracle_crossfit_whittaker_baseline_matrix(
    const double *rows, size_t m, size_t n, double lam, double *out)
{
    if (rows == NULL || out == NULL) return -1;
    for (size_t r = 0; r < m; ++r) {
        int st = oracle_crossfit_whittaker_baseline(rows + r * n, n, lam, out + r * n);
        if (st != 0) return -1;
    }
    return 0;
}

/* Exposes the data-independent factor so tests can check the GPU's head/steady/tail tables. */
int oracle_whittaker_factor(size_t n, int parity, double lam, double *d, double *l1, double *l2)
{
    if (n < 3) return -2;
    penta_factor f = { d, l1, l2 };
    factor_system(&f, n, parity, lam);
    return 0;
}

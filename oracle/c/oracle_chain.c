/*
 * ORACLE (test infrastructure only -- never linked into or called by the product path).
 *
 * CPU restatement of the reference's two-state chain DP.
 * Follows /root/reference/rocco/_chain_dp.c:109-186 (forward sweep with
 * (value, count) lexicographic tie-break, terminal choice, back-pointer chase).
 * Arithmetic order is kept identical so results are bit-for-bit those of the
 * reference; pinned against oracle/_ref (the reference compiled here) by
 * tests/test_oracle_pin.py.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

/* "a beats b": strictly larger value, or equal value with strictly fewer selected bins
 * (_chain_dp.c:133-134, 147-148, 167-168). */
static inline int beats(double va, long ka, double vb, long kb)
{
    return (va > vb) || (va == vb && ka < kb);
}

/* returns 0 ok, -1 alloc failure, -2 invalid input */
int oracle_solve_penalized_chain(
    const double *score, const double *cost, size_t n, double penalty,
    uint8_t *mask_out, double *value_out, long *count_out)
{
    if (score == NULL || mask_out == NULL || n == 0) return -2;
    if (n > 1 && cost == NULL) return -2;

    /* from0[i] = predecessor state of state 0 at bin i; from1[i] likewise for state 1 */
    uint8_t *from0 = (uint8_t *)calloc(n, 1);
    uint8_t *from1 = (uint8_t *)calloc(n, 1);
    if (!from0 || !from1) { free(from0); free(from1); return -1; }

    double v_off = 0.0;                 /* best value ending unselected */
    long k_off = 0;
    double v_on = score[0] - penalty;   /* best value ending selected   */
    long k_on = 1;

    for (size_t i = 1; i < n; ++i) {
        const double c = cost[i - 1];
        /* candidates, evaluated in the reference's order of operations */
        const double off_keep = v_off;
        const double off_leave = v_on - c;
        const double on_keep = v_on + score[i] - penalty;
        const double on_enter = v_off - c + score[i] - penalty;
        double nv_off, nv_on;
        long nk_off, nk_on;

        if (beats(off_leave, k_on, off_keep, k_off)) {
            nv_off = off_leave; nk_off = k_on; from0[i] = 1;
        } else {
            nv_off = off_keep; nk_off = k_off; from0[i] = 0;
        }
        if (beats(on_enter, k_off + 1, on_keep, k_on + 1)) {
            nv_on = on_enter; nk_on = k_off + 1; from1[i] = 0;
        } else {
            nv_on = on_keep; nk_on = k_on + 1; from1[i] = 1;
        }
        v_off = nv_off; k_off = nk_off;
        v_on = nv_on; k_on = nk_on;
    }

    int st;
    if (beats(v_on, k_on, v_off, k_off)) {
        st = 1; *value_out = v_on; *count_out = k_on;
    } else {
        st = 0; *value_out = v_off; *count_out = k_off;
    }
    mask_out[n - 1] = (uint8_t)st;
    for (size_t i = n - 1; i > 0; --i) {
        st = st ? from1[i] : from0[i];
        mask_out[i - 1] = (uint8_t)st;
    }
    free(from0);
    free(from1);
    return 0;
}

/* ORACLE -- test infrastructure only.
 *
 * Plain-C restatement of the per-read coverage loop of the reference's alignment counter,
 * /root/reference/rocco/native/ccounts_backend.c:2416-2574 (filters, fragment inference, strand shifts, extension,
 * clipping to the region, one-read-per-bin, delta buffer + float prefix sum), taking the fields htslib would
 * decode as plain arrays.  BAM decoding itself is out of scope.
 *
 * Parity pin: htslib cannot be built here, so this loop is NOT pinned against the compiled reference ("parity
 * unpinned" for these lines); everything downstream of it (readtracks.py:455-518, 590-633) is pinned by running the
 * reference's own Python over this counter (tests/golden/make_golden_assembly.py).
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

#define BAM_FPROPER_PAIR 2
#define BAM_FMUNMAP 8
#define BAM_FREVERSE 16
#define BAM_FREAD2 128

typedef struct {
    int32_t flag_include, flag_exclude, min_mapping_quality;
    int32_t paired_end_mode, one_read_per_bin;
    int64_t read_length, min_template_length, max_insert_size;
    int64_t shift_forward, shift_reverse, extend_bp;
} oracle_count_options;

int oracle_count_alignment_region(const int64_t *pos, const int64_t *endpos, const uint16_t *flag, const uint8_t *mapq,
                                  const int64_t *isize, const uint8_t *mate_same_tid, size_t n_reads,
                                  const oracle_count_options *opt, int64_t start, int64_t end, int64_t step,
                                  float *count_buffer, size_t count_len)
{
    float *delta = (float *)calloc(count_len + 1U, sizeof(float));
    if (!delta) return -1;
    const int64_t min_tlen = opt->min_template_length >= 0 ? opt->min_template_length : opt->read_length;
    for (size_t r = 0; r < n_reads; ++r) {
        int64_t adj_start, adj_end;
        if (opt->flag_include > 0 && (flag[r] & opt->flag_include) != opt->flag_include) continue;
        if ((flag[r] & opt->flag_exclude) != 0) continue;
        if ((int32_t)mapq[r] < opt->min_mapping_quality) continue;
        const int64_t read_start = pos[r], read_end = endpos[r];
        if (opt->paired_end_mode > 0) {
            if ((flag[r] & BAM_FPROPER_PAIR) == 0) continue;
            if ((flag[r] & BAM_FREAD2) != 0) continue;
            if ((flag[r] & BAM_FMUNMAP) != 0 || !mate_same_tid[r]) continue;
            const int64_t tlen = isize[r];
            const int64_t atlen = tlen >= 0 ? tlen : -tlen;
            if (atlen == 0 || atlen < min_tlen) continue;
            if (opt->max_insert_size > 0 && atlen > opt->max_insert_size) continue;
            if (tlen >= 0) { adj_start = read_start; adj_end = read_start + atlen; }
            else { adj_end = read_end; adj_start = adj_end - atlen; }
            if ((flag[r] & BAM_FREVERSE) == 0) { adj_start += opt->shift_forward; adj_end += opt->shift_forward; }
            else { adj_start -= opt->shift_reverse; adj_end -= opt->shift_reverse; }
        } else if ((flag[r] & BAM_FREVERSE) == 0) {
            const int64_t five = read_start + opt->shift_forward;
            if (opt->extend_bp > 0) { adj_start = five; adj_end = five + opt->extend_bp; }
            else { adj_start = read_start + opt->shift_forward; adj_end = read_end + opt->shift_forward; }
        } else {
            const int64_t five = (read_end - 1) - opt->shift_reverse;
            if (opt->extend_bp > 0) { adj_end = five + 1; adj_start = adj_end - opt->extend_bp; }
            else { adj_start = read_start - opt->shift_reverse; adj_end = read_end - opt->shift_reverse; }
        }
        if (adj_end <= start || adj_start >= end) continue;
        if (adj_start < start) adj_start = start;
        if (adj_end > end) adj_end = end;
        if (opt->one_read_per_bin) {
            const int64_t mid = (adj_start + adj_end) / 2;
            const size_t idx = (size_t)((mid - start) / step);
            if (idx < count_len) count_buffer[idx] += 1.0f;
            continue;
        }
        size_t i0 = (size_t)((adj_start - start) / step), i1 = (size_t)(((adj_end - 1) - start) / step);
        if (i0 >= count_len) continue;
        if (i1 >= count_len) i1 = count_len - 1U;
        if (i0 > i1) continue;
        delta[i0] += 1.0f;
        delta[i1 + 1U] -= 1.0f;
    }
    float run = 0.0f;
    for (size_t i = 0; i < count_len; ++i) { run += delta[i]; count_buffer[i] += run; }
    free(delta);
    return 0;
}

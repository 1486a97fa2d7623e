/*
 * ref_batch.c — ORACLE / CPU-BASELINE SHIM. TEST INFRASTRUCTURE ONLY.
 *
 * Compiled together with the reference's UNMODIFIED src/ssw.c (read in place from /root/reference by
 * oracle/Makefile, never copied) into oracle/_ref/libssw_ref.so.  It drives the reference's public
 * entry points in exactly the per-alignment call pattern of src/ssw_wrap.py:186-224
 * (ssw_init(score_size=2) -> ssw_align(flag=1, filters=0, filterd=0, maskLen) -> init_destroy ->
 * align_destroy) over a batch, so that (i) golden vectors come from the real reference and
 * (ii) the "ssw.c-only ceiling" CPU baseline of SURVEY.md §8d can be timed without Python overhead.
 */
#include <stdint.h>
#include <stdlib.h>
#include "ssw.h"

/* out: npairs x 7 int32 = score, ref_begin, ref_end, query_begin, query_end, score2, ref_end2.
 * cigar_out (optional): npairs x cigar_cap uint32, cigar_len (optional): npairs int32. */
int ref_align_batch(const int8_t *qbuf, const int64_t *qoff, const int8_t *tbuf, const int64_t *toff,
                    const int32_t *qidx, const int32_t *tidx, int64_t npairs,
                    const int8_t *mat, int nmat, int go, int ge, int32_t *out,
                    uint32_t *cigar_out, int32_t *cigar_len, int cigar_cap)
{
    for (int64_t k = 0; k < npairs; ++k) {
        int qi = qidx[k], ti = tidx[k];
        int m = (int)(qoff[qi + 1] - qoff[qi]);
        int n = (int)(toff[ti + 1] - toff[ti]);
        int mask = m > 30 ? m / 2 : 15;
        s_profile *p = ssw_init(qbuf + qoff[qi], m, mat, nmat, 2);
        s_align *a = ssw_align(p, tbuf + toff[ti], n, (uint8_t)go, (uint8_t)ge, 1, 0, 0, mask);
        int32_t *o = out + k * 7;
        if (!a) { o[0] = -1; o[1] = o[2] = o[3] = o[4] = o[5] = o[6] = -1; init_destroy(p); continue; }
        o[0] = a->score1; o[1] = a->ref_begin1; o[2] = a->ref_end1; o[3] = a->read_begin1;
        o[4] = a->read_end1; o[5] = a->score2; o[6] = a->ref_end2;
        if (cigar_out && cigar_len) {
            int L = a->cigarLen < cigar_cap ? a->cigarLen : cigar_cap;
            for (int c = 0; c < L; ++c) cigar_out[k * cigar_cap + c] = a->cigar[c];
            cigar_len[k] = a->cigarLen;
        }
        init_destroy(p);
        align_destroy(a);
    }
    return 0;
}

/*
 * sw_oracle.c — ORACLE. TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, scalar restatement of what the reference's striped Smith-Waterman returns *as tredparse
 * calls it* (flag=1, filters=0, filterd=0).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this; the product (tredparse_b200/) never does.
 *
 * Parity status: PINNED.  tests/test_oracle_sw.py checks every field against the reference's own
 * src/ssw.c compiled unmodified into oracle/_ref/libssw_ref.so (make ref) on the 24,500 real
 * (read, template) pairs of the two reference fixtures plus random / adversarial pairs, and against
 * the committed golden vectors in tests/golden/ (generated from that same library).
 *
 * What is restated (citations are into /root/reference/):
 *   recurrence, saturating E/F floors ............ src/ssw.c:203-232 (byte), 441-463 (word)
 *   lazy-F is value-neutral (I-then-D == D-then-I)  src/ssw.c:238-270, 467-479   
 *   forward tie-breaks (first column, min row) .... src/ssw.c:272-289, 299-308 / 482-496, 503-512
 *   reverse pass, terminate at score1 ............. src/ssw.c:178-183, 292-294, 839-851
 *   second-best (score2 / ref_end2) ............... src/ssw.c:315-340 (byte), 519-542 (word)
 *   byte->word switch at score+bias >= 255 ........ src/ssw.c:283, 807-810
 *   banded traceback -> CIGAR ..................... src/ssw.c:549-736
 *   post filter + read classification ............. src/ssw_wrap.py:213-220, tredparse/bam_parser.py:102-182
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t score;
    int32_t ref_begin;
    int32_t ref_end;
    int32_t query_begin;
    int32_t query_end;
    int32_t score2;
    int32_t ref_end2;
} tro_align_t;

static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }

/* One directional pass.  Columns are visited i = begin, begin+step, ... ; rows j = 0..m-1.
 * terminate < 0: never terminate.  colmax_out (optional, length n) receives per-column maxima of
 * the columns that were visited (others stay 0, like calloc in src/ssw.c:149).
 * end_ref_init mirrors src/ssw.c:145 (byte: -1) / :387 (word: 0). */
static void sw_pass(const int8_t *q, int m, int lanes, const int8_t *t, int n, const int8_t *mat, int nmat,
                    int go, int ge, int reverse_cols, int terminate, int end_ref_init,
                    int *o_score, int *o_end_ref, int *o_end_read, int32_t *colmax_out)
{
    /* The striped kernels round the query up to a whole number of SIMD vectors and give the padding
     * rows a substitution score of 0 (src/ssw.c:108, 363: `j >= readLen ? bias : ...`).  Those ghost
     * rows can never win the maximum (they only copy real values down the diagonal) and are excluded
     * from end_read, but they do show up in the per-column maxima behind score2 / ref_end2. */
    int mpad = ((m + lanes - 1) / lanes) * lanes;
    if (mpad < 1) mpad = 1;
    int32_t *H = (int32_t *)calloc((size_t)mpad, sizeof(int32_t));
    int32_t *E = (int32_t *)calloc((size_t)mpad, sizeof(int32_t));
    int32_t *Hmax = (int32_t *)calloc((size_t)mpad, sizeof(int32_t));
    int max = 0, end_ref = end_ref_init, end_read = m - 1;
    int begin = 0, end = n, step = 1;
    if (reverse_cols) { begin = n - 1; end = -1; step = -1; }
    for (int i = begin; i != end; i += step) {
        const int8_t *row = mat + (int)t[i] * nmat;
        int F = 0, hdiag = 0, colmax = 0;
        for (int j = 0; j < mpad; ++j) {
            int hprev = H[j];
            int h = hdiag + (j < m ? row[(int)q[j]] : 0);
            h = imax(h, E[j]);
            h = imax(h, F);
            h = imax(h, 0);
            H[j] = h;
            int hg = h - go;
            E[j] = imax(imax(E[j] - ge, hg), 0);
            F = imax(imax(F - ge, hg), 0);
            hdiag = hprev;
            colmax = imax(colmax, h);
        }
        if (colmax > max) {            /* strict: first column wins ties */
            max = colmax;
            end_ref = i;
            memcpy(Hmax, H, (size_t)mpad * sizeof(int32_t));
        }
        if (colmax_out) colmax_out[i] = colmax;
        if (terminate >= 0 && colmax == terminate) break;
    }
    /* smallest row index holding the maximum in the winning column; starts at m-1 and only moves up
     * (src/ssw.c:299-308) — also reproduces the all-zero degenerate case (end_read = 0). */
    for (int j = 0; j < m; ++j) {
        if (Hmax[j] == max && j < end_read) end_read = j;
    }
    *o_score = max; *o_end_ref = end_ref; *o_end_read = end_read;
    free(H); free(E); free(Hmax);
}

/* Full ssw_align(flag=1) restatement.  mat is nmat x nmat (tredparse: 5x5), bias = |min(mat)|. */
int tro_align(const int8_t *q, int m, const int8_t *t, int n, const int8_t *mat, int nmat,
              int go, int ge, int mask_len, tro_align_t *out)
{
    int bias = 0;
    for (int i = 0; i < nmat * nmat; ++i) if (mat[i] < bias) bias = mat[i];
    bias = -bias;
    int32_t *colmax = (int32_t *)calloc((size_t)(n > 0 ? n : 1), sizeof(int32_t));
    int score, end_ref, end_read;
    /* Which kernel would the reference end in?  The 8-bit one (16 lanes) unless the score saturates
     * it (score + bias >= 255, src/ssw.c:283,807-810), then the 16-bit one (8 lanes). */
    sw_pass(q, m, 16, t, n, mat, nmat, go, ge, 0, -1, -1, &score, &end_ref, &end_read, colmax);
    int word = (score + bias >= 255);
    if (word) {
        memset(colmax, 0, (size_t)(n > 0 ? n : 1) * sizeof(int32_t));
        sw_pass(q, m, 8, t, n, mat, nmat, go, ge, 0, -1, 0, &score, &end_ref, &end_read, colmax);
    }
    out->score = score; out->ref_end = end_ref; out->query_end = end_read;

    /* second best (ignored by tredparse, reported for completeness of the legacy ABI) */
    out->score2 = 0; out->ref_end2 = -1;
    if (mask_len >= 15) {
        int s2 = 0, r2 = 0;
        int edge = (end_ref - mask_len) > 0 ? (end_ref - mask_len) : 0;
        for (int i = 0; i < edge; ++i) if (colmax[i] > s2) { s2 = colmax[i]; r2 = i; }
        edge = (end_ref + mask_len) > n ? n : (end_ref + mask_len);
        for (int i = edge + (word ? 0 : 1); i < n; ++i) if (colmax[i] > s2) { s2 = colmax[i]; r2 = i; }
        out->score2 = s2; out->ref_end2 = r2;
    }
    free(colmax);

    if (end_ref < 0) {                              /* score == 0, byte path: nothing aligned */
        out->ref_begin = -1; out->query_begin = 0;
        return 0;
    }
    /* reverse pass on q[0..end_read] reversed, template columns end_ref..0 */
    int m2 = end_read + 1, n2 = end_ref + 1;
    int8_t *qr = (int8_t *)malloc((size_t)m2);
    for (int j = 0; j < m2; ++j) qr[j] = q[end_read - j];
    int s_r, ref_r, read_r;
    sw_pass(qr, m2, word ? 8 : 16, t, n2, mat, nmat, go, ge, 1, score, word ? 0 : -1, &s_r, &ref_r, &read_r, NULL);
    free(qr);
    out->ref_begin = ref_r;
    out->query_begin = end_read - read_r;
    return 0;
}

/* pairs: for k in [0,npairs): query qidx[k], template tidx[k]; sequences are flat int8 code buffers
 * with offset arrays (qoff has nq+1 entries, toff has nt+1). out is npairs x 7 int32. */
int tro_align_batch(const int8_t *qbuf, const int64_t *qoff, const int8_t *tbuf, const int64_t *toff,
                    const int32_t *qidx, const int32_t *tidx, int64_t npairs,
                    const int8_t *mat, int nmat, int go, int ge, int32_t *out)
{
    for (int64_t k = 0; k < npairs; ++k) {
        int qi = qidx[k], ti = tidx[k];
        int m = (int)(qoff[qi + 1] - qoff[qi]);
        int n = (int)(toff[ti + 1] - toff[ti]);
        int mask = m > 30 ? m / 2 : 15;             /* src/ssw_wrap.py:198-201 */
        tro_align_t r;
        tro_align(qbuf + qoff[qi], m, tbuf + toff[ti], n, mat, nmat, go, ge, mask, &r);
        int32_t *o = out + k * 7;
        o[0] = r.score; o[1] = r.ref_begin; o[2] = r.ref_end; o[3] = r.query_begin;
        o[4] = r.query_end; o[5] = r.score2; o[6] = r.ref_end2;
    }
    return 0;
}

/* Tags (tredparse/bam_parser.py:157-168).  0 = dropped / filtered. */
enum { TRO_NONE = 0, TRO_FULL = 1, TRO_PREF = 2, TRO_POST = 3, TRO_REPT = 4, TRO_HANG = 5 };
#define TRO_FLANKMATCH 9

/* Post-filter (src/ssw_wrap.py:213-220) + classification (tredparse/bam_parser.py:133-168) of one
 * (read, template) alignment.  m = read length, n = template length, u = repeat units of the
 * template, max_units_eff = ceil(m/period) if clip else ceil(READLEN/period). */
int tro_classify(int score, int rb, int re, int qb, int qe, int m, int n, int u, int period,
                 int max_units_eff)
{
    int min_len = imin(m, n) / 2;
    int min_score = imax(min_len, 30);
    if (!(score >= min_score && (qe - qb + 1) >= min_len)) return TRO_NONE;
    int prefix_read = rb < TRO_FLANKMATCH;
    int suffix_read = re > n - TRO_FLANKMATCH - 1;
    int aL = rb, aR = n - re - 1, bL = qb, bR = m - qe - 1;
    int hang = imin(imin(aR + bL, aL + bR), imin(aL + aR, bL + bR));
    if (hang >= TRO_FLANKMATCH) return TRO_HANG;
    if (prefix_read) return suffix_read ? TRO_FULL : TRO_PREF;
    if (suffix_read) return TRO_POST;
    if (u >= max_units_eff - 1 && u * period <= m) return TRO_REPT;
    return TRO_NONE;
}

/* One read against a locus' template family in DB order (units ascending, forward template before its
 * reverse complement; tredparse/bam_parser.py:84-100) and the per-read arg-max of (score, -units),
 * first seen wins (bam_parser.py:174).  Templates are supplied pre-built: tbuf/toff hold 2*max_units
 * sequences in DB order.  Returns tag (0 if the read yields nothing); *o_score, *o_h filled.
 * pair_out (optional, 2*max_units x 7) receives every alignment. */
int tro_classify_read(const int8_t *q, int m, const int8_t *tbuf, const int64_t *toff, int max_units,
                      int period, int max_units_eff, const int8_t *mat, int nmat, int go, int ge,
                      int *o_score, int *o_h, int32_t *pair_out)
{
    int best_tag = TRO_NONE, best_score = -1, best_h = 0;
    int mask = m > 30 ? m / 2 : 15;
    for (int k = 0; k < 2 * max_units; ++k) {
        int u = k / 2 + 1;
        int n = (int)(toff[k + 1] - toff[k]);
        tro_align_t r;
        tro_align(q, m, tbuf + toff[k], n, mat, nmat, go, ge, mask, &r);
        if (pair_out) {
            int32_t *o = pair_out + (int64_t)k * 7;
            o[0] = r.score; o[1] = r.ref_begin; o[2] = r.ref_end; o[3] = r.query_begin;
            o[4] = r.query_end; o[5] = r.score2; o[6] = r.ref_end2;
        }
        int tag = tro_classify(r.score, r.ref_begin, r.ref_end, r.query_begin, r.query_end,
                               m, n, u, period, max_units_eff);
        if (tag == TRO_NONE) continue;
        /* max over key (score, -units): strictly better score, or same score with fewer units;
         * equal key keeps the first seen. */
        if (r.score > best_score || (r.score == best_score && u < best_h)) {
            best_score = r.score; best_h = u; best_tag = tag;
        }
    }
    *o_score = best_score; *o_h = best_h;
    return best_tag;
}

/* ------------------------------------------------------------------------------------------------
 * CIGAR: restatement of banded_sw (src/ssw.c:549-736) on the sub-sequences
 * ref[ref_begin..ref_end], read[query_begin..query_end], band = |refLen-readLen|+1 doubled until
 * the score is reproduced.  Output words are len<<4|op with M/I/D = 0/1/2 (src/ssw.h:132-170).
 * Returns the number of words written (<= cap), or -1 on traceback error / overflow.
 * ---------------------------------------------------------------------------------------------- */
static inline int band_u(int w, int i, int j) { int x = i - w; x = x > 0 ? x : 0; return j - x + 1; }
static inline int band_d(int w, int i, int j, int p) { int x = i - w; x = x > 0 ? x : 0; x = j - x; return x * 3 + p; }

int tro_cigar(const int8_t *ref, const int8_t *read, int refLen, int readLen, int score,
              int go, int ge, const int8_t *mat, int nmat, uint32_t *out, int cap)
{
    int band_width = abs(refLen - readLen) + 1;
    int width, width_d, max = 0;
    int32_t *h_b = NULL, *e_b = NULL, *h_c = NULL;
    int8_t *direction = NULL, *direction_line;
    int iter = 0;
    do {
        width = band_width * 2 + 3; width_d = band_width * 2 + 1;
        h_b = (int32_t *)realloc(h_b, (size_t)(width + 1) * sizeof(int32_t));
        e_b = (int32_t *)realloc(e_b, (size_t)(width + 1) * sizeof(int32_t));
        h_c = (int32_t *)realloc(h_c, (size_t)(width + 1) * sizeof(int32_t));
        direction = (int8_t *)realloc(direction, (size_t)width_d * readLen * 3 + 16);
        if (++iter > 40) { free(h_b); free(e_b); free(h_c); free(direction); return -1; }
        max = 0;
        for (int j = 1; j < width - 1; ++j) h_b[j] = 0;
        for (int i = 0; i < readLen; ++i) {
            int beg = 0, end = refLen - 1, u = 0, edge, j, f;
            j = i - band_width; beg = beg > j ? beg : j;
            j = i + band_width; end = end < j ? end : j;
            edge = end + 1 < width - 1 ? end + 1 : width - 1;
            f = h_b[0] = e_b[0] = h_b[edge] = e_b[edge] = h_c[0] = 0;
            direction_line = direction + (size_t)width_d * i * 3;
            for (j = beg; j <= end; ++j) {
                int b, e, e1, f1, d, de, df, dh, temp1, temp2;
                u = band_u(band_width, i, j); e = band_u(band_width, i - 1, j);
                b = band_u(band_width, i, j - 1); d = band_u(band_width, i - 1, j - 1);
                de = band_d(band_width, i, j, 0);
                df = band_d(band_width, i, j, 1);
                dh = band_d(band_width, i, j, 2);
                temp1 = i == 0 ? -go : h_b[e] - go;
                temp2 = i == 0 ? -ge : e_b[e] - ge;
                e_b[u] = temp1 > temp2 ? temp1 : temp2;
                direction_line[de] = temp1 > temp2 ? 3 : 2;
                temp1 = h_c[b] - go;
                temp2 = f - ge;
                f = temp1 > temp2 ? temp1 : temp2;
                direction_line[df] = temp1 > temp2 ? 5 : 4;
                e1 = e_b[u] > 0 ? e_b[u] : 0;
                f1 = f > 0 ? f : 0;
                temp1 = e1 > f1 ? e1 : f1;
                temp2 = h_b[d] + mat[(int)ref[j] * nmat + (int)read[i]];
                h_c[u] = temp1 > temp2 ? temp1 : temp2;
                if (h_c[u] > max) max = h_c[u];
                if (temp1 <= temp2) direction_line[dh] = 1;
                else direction_line[dh] = e1 > f1 ? direction_line[de] : direction_line[df];
            }
            for (j = 1; j <= u; ++j) h_b[j] = h_c[j];
        }
        band_width *= 2;
    } while (max < score);
    band_width /= 2;
    width_d = band_width * 2 + 1;

    /* trace back from the bottom-right corner */
    int i = readLen - 1, j = refLen - 1, e = 0, l = 0, temp2 = 2, n_out = 0, err = 0;
    char op = 'M', prev_op = 'M';
    uint32_t *c = (uint32_t *)malloc((size_t)(readLen + refLen + 4) * sizeof(uint32_t));
    direction_line = direction + (size_t)width_d * (readLen - 1) * 3;
    static const int opcode[256] = { ['M'] = 0, ['I'] = 1, ['D'] = 2 };
    while (i > 0) {
        int temp1 = band_d(band_width, i, j, temp2);
        switch (direction_line[temp1]) {
            case 1: --i; --j; temp2 = 2; direction_line -= width_d * 3; op = 'M'; break;
            case 2: --i; temp2 = 0; direction_line -= width_d * 3; op = 'I'; break;
            case 3: --i; temp2 = 2; direction_line -= width_d * 3; op = 'I'; break;
            case 4: --j; temp2 = 1; op = 'D'; break;
            case 5: --j; temp2 = 2; op = 'D'; break;
            default: err = 1; break;
        }
        if (err) break;
        if (op == prev_op) ++e;
        else {
            c[l++] = ((uint32_t)e << 4) | (uint32_t)opcode[(unsigned char)prev_op];
            prev_op = op; e = 1;
        }
    }
    if (!err) {
        if (op == 'M') c[l++] = ((uint32_t)(e + 1) << 4) | 0u;
        else { c[l++] = ((uint32_t)e << 4) | (uint32_t)opcode[(unsigned char)op]; c[l++] = (1u << 4) | 0u; }
        if (l <= cap) { for (int k = 0; k < l; ++k) out[k] = c[l - 1 - k]; n_out = l; }
        else n_out = -1;
    } else n_out = -1;
    free(c); free(h_b); free(e_b); free(h_c); free(direction);
    return n_out;
}

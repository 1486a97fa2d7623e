// sw_family.cu — the production Smith-Waterman path: every read of a locus against the whole template
// family of that locus (prefix + repeat*u + suffix and reverse complements, u = 1..max_units), with the
// reference's post-filter, classification and per-read arg-max fused in.
//
// Replaces, for a whole batch of (sample, locus) problems, the loop
//     for read: for (units, target, ssw) in db: ssw.align(...) ; classify ; max(res)
// of tredparse/bam_parser.py:123-182 over src/ssw_wrap.py:177-227 over src/ssw.c:780-871.
//
// Mapping: one warp = 32 reads of the same family, one thread = one read.  Every lane walks the same
// template columns at the same time, so template bases are warp-uniform; query bases are per lane.
//
// Phase 1 (scores of all 2*max_units templates):
//   * the templates of a family are nested — prefix + repeat*u is a prefix of prefix + repeat*(u+1) —
//     so the DP columns of the shared part are computed once and only the suffix columns are "forked"
//     per u.  A strip = the P columns of repeat unit u followed by the Ls suffix columns of template u,
//     all in registers (previous-row H and running F per column); query rows stream through; the H/E
//     column leaving the repeat unit goes to a per-lane boundary column in shared memory.
//   * the forward template and its reverse complement have identical shape, so both are computed in one
//     pass as the two int16 halves of packed registers with DPX instructions
//     (VIADDMNMX.S16x2.RELU, VIMNMX.S16x2); substitution scores for both halves come from one PRMT.
//   ssw's outputs only depend on column maxima / first-maximum positions, so this is exact.
// Phase 2 (positions of the winning template only): candidates are visited in the reference's arg-max
//   order (score desc, units asc, forward before reverse complement); for each, the exact end / begin
//   coordinates come from the scalar sweeps of sw_sweep.cuh (forward locate + reverse pass), then the
//   reference's filter + tag rules; the first candidate that yields a tag is the read's result — the
//   same result as classifying all 2*max_units alignments and taking max(key=(score, -units)).
#include "internal.cuh"
#include "sw_sweep.cuh"

namespace {

constexpr int FAM_W2 = 16;         // strip width of the scalar phase-2 sweeps
constexpr int FLANK = 18;          // flank length of the fast path (all catalogue loci)

struct FamilySmem {
    int8_t prefix[32], suffix[32], repeat[32];
    int Lp, Ls, P, U, clip;
};

struct ClassifyParams {
    const int8_t *rbuf; const int64_t *roff;
    const int32_t *order;          // read indices grouped by family
    const int32_t *fam_start;      // [nfam+1] offsets into order
    const int32_t *chunk_start;    // [nfam+1] offsets into the item (warp-chunk) list
    const tredsw_family *families;
    int nfamilies;
    int go, ge;
    int max_rows;
    int allow_fast;                // scores fit the 8-bit boundary column (max_read_len * match < 256)
    int32_t *out;
    unsigned long long *stats;     // 4 counters
};

__constant__ int8_t c_fmat25[25];

__device__ __forceinline__ int fam_code(const FamilySmem &F, int u, int strand, int n, int i) {
    int k = strand ? (n - 1 - i) : i;
    int c;
    if (k < F.Lp) c = F.prefix[k];
    else if (k < F.Lp + F.P * u) c = F.repeat[(k - F.Lp) % F.P];
    else c = F.suffix[k - F.Lp - F.P * u];
    return strand ? (c < 4 ? 3 - c : c) : c;
}

// ---------------------------------------------------------------------------------------------------
// Phase 1, generic: one scalar sweep per template.  Any flank length / period.  Used when the family
// does not fit the packed fast path, and as an on-device cross-check of it.
// ---------------------------------------------------------------------------------------------------
__device__ void phase1_generic(const FamilySmem &F, const SwLut *lut, const uint8_t *codes, int lane,
                               int m, int m_warp, uint32_t *bnd, uint16_t *scores, int go, int ge,
                               unsigned long long &cells) {
    auto rc = [&](int j) { return j < m ? (int)codes[j * 32 + lane] : SW_CODE_GHOST; };
    for (int rank = 0; rank < 2 * F.U; ++rank) {
        const int u = rank / 2 + 1, s = rank & 1;
        const int n = F.Lp + F.Ls + F.P * u;
        auto cc = [&](int i) { return fam_code(F, u, s, n, i); };
        int dc, dr;
        int score = sw_sweep<FAM_W2, 0, false>(m, m, n, rc, cc, lut, bnd + lane, 32, go, ge, 0, &dc, &dr, nullptr);
        scores[rank * 32 + lane] = (uint16_t)min(score, 65535);
        cells += (unsigned long long)m * n;
    }
}

// ---------------------------------------------------------------------------------------------------
// Phase 1, fast: nested strips, forward + reverse-complement packed as s16x2.
// Requires Lp == Ls == FLANK and scores < 256 (8-bit boundary column).
// Boundary word per row and lane: H_fwd | H_rc << 8 | E_fwd << 16 | E_rc << 24.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sel2(int tf, int tr) {   // PRMT selector: sign-extended bytes -> s16x2
    return (uint32_t)tf | ((uint32_t)(8 | tf) << 4) | ((uint32_t)tr << 8) | ((uint32_t)(8 | tr) << 12);
}

// columns [C0, C1) of a strip whose previous-row H and running F live in registers
template <int NC, int C0, int C1>
__device__ __forceinline__ void packed_row(const uint32_t (&sel)[NC], uint32_t (&Hrow)[NC], uint32_t (&Fv)[NC],
                                           uint32_t w0, uint32_t w1, uint32_t &hd, uint32_t &e, uint32_t &mx,
                                           uint32_t mgo2, uint32_t mge2) {
    constexpr uint32_t NEG = 0x80008000u;    // (-32768, -32768): max with it is the identity
#pragma unroll
    for (int c = C0; c < C1; ++c) {
        const uint32_t s = sw_prmt(w0, w1, sel[c]);             // s16x2 substitution scores
        const uint32_t t = __vmaxs2(e, Fv[c]);
        const uint32_t h = __viaddmax_s16x2_relu(hd, s, t);         // max(diag + s, E, F, 0)
        hd = Hrow[c];
        Hrow[c] = h;
        const uint32_t hgo = __viaddmax_s16x2(h, mgo2, NEG);        // h - go per half
        e = __viaddmax_s16x2_relu(e, mge2, hgo);                    // max(E - ge, h - go, 0)
        Fv[c] = __viaddmax_s16x2_relu(Fv[c], mge2, hgo);
        mx = __vmaxs2(mx, h);
    }
}

template <int P>
__device__ void phase1_packed(const FamilySmem &F, const SwLut *lut, const uint8_t *codes, int lane, int m,
                              int m_warp, uint32_t *bnd, uint16_t *scores, int go, int ge,
                              unsigned long long &cells) {
    constexpr int NC = P + FLANK;
    const uint32_t mgo2 = (uint32_t)((-go) & 0xffff) | ((uint32_t)((-go) & 0xffff) << 16);
    const uint32_t mge2 = (uint32_t)((-ge) & 0xffff) | ((uint32_t)((-ge) & 0xffff) << 16);
    // rc family: prefix' = rc(suffix), repeat' = rc(repeat), suffix' = rc(prefix)
    auto comp = [](int c) { return c < 4 ? 3 - c : c; };
    uint32_t sel[NC];
    uint32_t Hrow[NC], Fv[NC];
    uint32_t m_main = 0;          // running maximum over the shared (main) columns, per strand
    // ---- strip 0: the FLANK prefix columns (no fork) -------------------------------------------------
    {
        uint32_t selp[FLANK], Hp[FLANK], Fp[FLANK];
#pragma unroll
        for (int c = 0; c < FLANK; ++c) {
            selp[c] = sel2(F.prefix[c], comp(F.suffix[FLANK - 1 - c]));
            Hp[c] = 0; Fp[c] = 0;
        }
        for (int j = 0; j < m_warp; ++j) {
            const int code = j < m ? (int)codes[j * 32 + lane] : SW_CODE_GHOST;
            const uint32_t w0 = lut->w0[code], w1 = lut->w1[code];
            uint32_t hd = 0, e = 0;
            packed_row<FLANK, 0, FLANK>(selp, Hp, Fp, w0, w1, hd, e, m_main, mgo2, mge2);
            // H (bytes 0,2 of the packed halves) and E -> 8-bit boundary word
            bnd[j * 32 + lane] = sw_prmt(Hp[FLANK - 1], e, 0x6420);
        }
    }
    // selectors of one strip: P repeat columns then FLANK suffix columns
#pragma unroll
    for (int c = 0; c < P; ++c) sel[c] = sel2(F.repeat[c], comp(F.repeat[P - 1 - c]));
#pragma unroll
    for (int c = 0; c < FLANK; ++c) sel[P + c] = sel2(F.suffix[c], comp(F.prefix[FLANK - 1 - c]));
    // ---- strips 1..U -----------------------------------------------------------------------------------
    for (int u = 1; u <= F.U; ++u) {
#pragma unroll
        for (int c = 0; c < NC; ++c) { Hrow[c] = 0; Fv[c] = 0; }
        uint32_t m_suf = 0;
        uint32_t hin_prev = 0;
        for (int j = 0; j < m_warp; ++j) {
            const uint32_t b = bnd[j * 32 + lane];
            const uint32_t hin = sw_prmt(b, 0, 0x4140);      // bytes 0,1 -> halves
            uint32_t e = sw_prmt(b, 0, 0x4342);              // bytes 2,3 -> halves
            const int code = j < m ? (int)codes[j * 32 + lane] : SW_CODE_GHOST;
            const uint32_t w0 = lut->w0[code], w1 = lut->w1[code];
            uint32_t hd = hin_prev;
            hin_prev = hin;
            packed_row<NC, 0, P>(sel, Hrow, Fv, w0, w1, hd, e, m_main, mgo2, mge2);
            bnd[j * 32 + lane] = sw_prmt(Hrow[P - 1], e, 0x6420);
            packed_row<NC, P, NC>(sel, Hrow, Fv, w0, w1, hd, e, m_suf, mgo2, mge2);
        }
        const uint32_t best = __vmaxs2(m_main, m_suf);
        scores[(2 * (u - 1) + 0) * 32 + lane] = (uint16_t)(best & 0xffffu);
        scores[(2 * (u - 1) + 1) * 32 + lane] = (uint16_t)(best >> 16);
    }
    cells += (unsigned long long)m * 2ull * (unsigned long long)(FLANK + F.U * NC);
}

// ---------------------------------------------------------------------------------------------------
template <int FAST_P>     // 0 = generic phase 1
__global__ void __launch_bounds__(32) classify_kernel(ClassifyParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ SwLut lut;
    __shared__ FamilySmem F;
    const int lane = threadIdx.x;
    const int item = blockIdx.x;
    // item -> family (binary search over chunk_start)
    int lo = 0, hi = p.nfamilies;
    if (item >= p.chunk_start[p.nfamilies]) return;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (p.chunk_start[mid] <= item) lo = mid; else hi = mid;
    }
    const int f = lo;
    {
        const tredsw_family &g = p.families[f];
        if (lane < 32) { F.prefix[lane] = g.prefix[lane]; F.suffix[lane] = g.suffix[lane]; F.repeat[lane] = g.repeat[lane]; }
        if (lane == 0) { F.Lp = g.prefix_len; F.Ls = g.suffix_len; F.P = g.period; F.U = g.max_units; F.clip = g.clip; }
    }
    sw_build_lut(&lut, c_fmat25, lane, 32);
    __syncwarp();
    // Only families matching this instantiation are processed here (the host launches one
    // instantiation per shape class over the same item list).
    const bool fast_shape = p.allow_fast && (F.Lp == FLANK && F.Ls == FLANK && F.P >= 1 && F.P <= 12);
    if (FAST_P == 0) { if (fast_shape) return; }
    else { if (!(fast_shape && F.P == FAST_P)) return; }

    uint32_t *bnd = reinterpret_cast<uint32_t *>(smem_raw);                       // [max_rows][32]
    uint8_t *codes = reinterpret_cast<uint8_t *>(bnd + (size_t)p.max_rows * 32);  // [max_rows][32]
    uint16_t *scores = reinterpret_cast<uint16_t *>(codes + (size_t)p.max_rows * 32);  // [2U][32]

    const int idx = p.fam_start[f] + 32 * (item - p.chunk_start[f]) + lane;
    const bool valid = idx < p.fam_start[f + 1];
    const int r = valid ? p.order[idx] : -1;
    int m = 0;
    const int8_t *q = nullptr;
    if (valid) { q = p.rbuf + p.roff[r]; m = (int)(p.roff[r + 1] - p.roff[r]); }
    bool too_long = m > p.max_rows;
    if (too_long) m = 0;
    for (int j = 0; j < m; ++j) { int c = q[j]; codes[j * 32 + lane] = (uint8_t)((c < 0 || c > 4) ? 4 : c); }
    const int m_warp = __reduce_max_sync(0xffffffffu, m);
    __syncwarp();

    unsigned long long cells1 = 0, cells2 = 0;
    if (FAST_P == 0) phase1_generic(F, &lut, codes, lane, m, m_warp, bnd, scores, p.go, p.ge, cells1);
    else phase1_packed<(FAST_P == 0 ? 1 : FAST_P)>(F, &lut, codes, lane, m, m_warp, bnd, scores, p.go, p.ge, cells1);

    // ---- Phase 2: walk candidates in arg-max order until one yields a tag ------------------------------
    int tag = TREDSW_TAG_NONE, best_u = 0, best_score = -1, rb = -1, re = -1, qb = -1, qe = -1, best_rank = -1;
    if (valid && m > 0) {
        const int max_units_eff = F.clip ? (m + F.P - 1) / F.P : F.U;
        int last_score = 0x7fffffff, last_rank = -1;
        auto rc_f = [&](int j) { return (int)codes[j * 32 + lane]; };
        for (;;) {
            int cs = -1, cr = -1;
            for (int rank = 0; rank < 2 * F.U; ++rank) {
                const int sc = scores[rank * 32 + lane];
                const int n = F.Lp + F.Ls + F.P * (rank / 2 + 1);
                const int min_len = min(m, n) / 2;
                if (sc < max(min_len, 30)) continue;
                if (sc > last_score || (sc == last_score && rank <= last_rank)) continue;   // already tried
                if (sc > cs) { cs = sc; cr = rank; }
            }
            if (cr < 0) break;
            last_score = cs; last_rank = cr;
            const int u = cr / 2 + 1, s = cr & 1;
            const int n = F.Lp + F.Ls + F.P * u;
            auto cc_f = [&](int i) { return fam_code(F, u, s, n, i); };
            int end_ref, end_read;
            sw_sweep<FAM_W2, 1, false>(m, m, n, rc_f, cc_f, &lut, bnd + lane, 32, p.go, p.ge, cs, &end_ref, &end_read, nullptr);
            cells2 += (unsigned long long)m * min(n, (end_ref / FAM_W2 + 1) * FAM_W2);
            if (end_ref < 0) continue;   // cannot happen: the score was produced by this very template
            auto rc_r = [&](int j) { return (int)codes[(end_read - j) * 32 + lane]; };
            auto cc_r = [&](int i) { return fam_code(F, u, s, n, end_ref - i); };
            int ci, rj;
            sw_sweep<FAM_W2, 1, false>(end_read + 1, end_read + 1, end_ref + 1, rc_r, cc_r, &lut, bnd + lane, 32, p.go,
                                       p.ge, cs, &ci, &rj, nullptr);
            cells2 += (unsigned long long)(end_read + 1) * min(end_ref + 1, (ci / FAM_W2 + 1) * FAM_W2);
            const int c_rb = end_ref - ci, c_qb = end_read - rj;
            const int t = sw_classify(cs, c_rb, end_ref, c_qb, end_read, m, n, u, F.P, max_units_eff);
            if (t != TREDSW_TAG_NONE) {
                tag = t; best_u = u; best_score = cs; rb = c_rb; re = end_ref; qb = c_qb; qe = end_read;
                best_rank = cr;
                break;
            }
        }
    }
    if (valid) {
        int32_t *o = p.out + (int64_t)r * 8;
        o[0] = too_long ? -1 : tag; o[1] = best_u; o[2] = best_score; o[3] = rb; o[4] = re; o[5] = qb; o[6] = qe;
        o[7] = best_rank;
    }
    if (p.stats) {
        unsigned long long alg = 0;
        if (valid && m > 0) {
            unsigned long long sum_n = 0;
            for (int u = 1; u <= F.U; ++u) sum_n += 2ull * (unsigned long long)(F.Lp + F.Ls + F.P * u);
            alg = (unsigned long long)m * sum_n;
        }
        unsigned long long c1 = cells1, c2 = cells2;
        for (int d = 16; d > 0; d >>= 1) {
            alg += __shfl_down_sync(0xffffffffu, alg, d);
            c1 += __shfl_down_sync(0xffffffffu, c1, d);
            c2 += __shfl_down_sync(0xffffffffu, c2, d);
        }
        unsigned nal = __reduce_add_sync(0xffffffffu, (valid && m > 0) ? (unsigned)(2 * F.U) : 0u);
        if (lane == 0) {
            atomicAdd(&p.stats[0], alg); atomicAdd(&p.stats[1], c1); atomicAdd(&p.stats[2], c2);
            atomicAdd(&p.stats[3], (unsigned long long)nal);
        }
    }
}

// ---- grouping of reads by family on the device (counting sort) ------------------------------------
__global__ void fam_count_kernel(const int32_t *read_family, int nreads, int nfam, int32_t *count) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < nreads) {
        int f = read_family[r];
        if (f >= 0 && f < nfam) atomicAdd(&count[f], 1);
    }
}
// single block: exclusive scans -> fam_start, chunk_start; cursor := fam_start
__global__ void fam_scan_kernel(const int32_t *count, int nfam, int32_t *fam_start, int32_t *chunk_start,
                                int32_t *cursor) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int a = 0, b = 0;
        for (int f = 0; f < nfam; ++f) {
            fam_start[f] = a; chunk_start[f] = b; cursor[f] = a;
            a += count[f]; b += (count[f] + 31) / 32;
        }
        fam_start[nfam] = a; chunk_start[nfam] = b;
    }
}
__global__ void fam_scatter_kernel(const int32_t *read_family, int nreads, int nfam, int32_t *cursor,
                                   int32_t *order, int32_t *out) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < nreads) {
        int f = read_family[r];
        if (f >= 0 && f < nfam) order[atomicAdd(&cursor[f], 1)] = r;
        else { int32_t *o = out + (int64_t)r * 8; o[0] = -1; o[1] = o[2] = o[3] = o[4] = o[5] = o[6] = o[7] = -1; }
    }
}

template <int FAST_P>
int launch_classify(tredsw_ctx *ctx, const ClassifyParams &p, int nitems_bound, size_t smem) {
    if (smem > 48 * 1024)
        CUDA_TRY(cudaFuncSetAttribute(classify_kernel<FAST_P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    classify_kernel<FAST_P><<<nitems_bound, 32, smem, ctx->stream>>>(p);
    CUDA_TRY(cudaGetLastError());
    ctx->launches += 1;
    return TREDSW_OK;
}

}  // namespace

int tredsw_internal_classify(tredsw_ctx *ctx, const int8_t *d_rbuf, const int64_t *d_roff, int nreads,
                             const int32_t *d_read_family, const tredsw_family *d_families,
                             const tredsw_family *h_families, int nfamilies, int max_m,
                             const int8_t *mat25, int gap_open, int gap_extend, int32_t *d_work, int32_t *d_out,
                             unsigned long long *d_stats) {
    int max_u = 0; unsigned pmask = 0; bool need_generic = false; int max_match = 0;
    for (int i = 0; i < 25; ++i) if (mat25[i] > max_match) max_match = mat25[i];
    const bool allow_fast = (long long)max_m * max_match < 256;
    for (int f = 0; f < nfamilies; ++f) {
        const tredsw_family &g = h_families[f];
        if (g.prefix_len < 1 || g.prefix_len > 32 || g.suffix_len < 1 || g.suffix_len > 32 || g.period < 1 ||
            g.period > 32 || g.max_units < 1 || g.max_units > 4096) { tredsw_set_error("family %d out of range", f); return TREDSW_ERR_ARG; }
        if (g.max_units > max_u) max_u = g.max_units;
        bool fast = allow_fast && g.prefix_len == FLANK && g.suffix_len == FLANK && g.period <= 12;
        if (fast) pmask |= 1u << g.period; else need_generic = true;
    }
    const int max_rows = max_m > 0 ? max_m : 1;
    const size_t smem = (size_t)max_rows * 32 * 4 + (size_t)max_rows * 32 + (size_t)2 * max_u * 32 * 2 + 64;
    if (smem > ctx->smem_optin) { tredsw_set_error("reads too long for the shared-memory boundary column (%zu B)", smem); return TREDSW_ERR_UNSUPPORTED; }
    ClassifyParams p{};
    p.rbuf = d_rbuf; p.roff = d_roff; p.families = d_families; p.out = d_out;
    int32_t *w = d_work;
    int32_t *d_count = w, *d_fam_start = w + nfamilies, *d_chunk_start = d_fam_start + nfamilies + 1,
            *d_cursor = d_chunk_start + nfamilies + 1, *d_order = d_cursor + nfamilies;
    CUDA_TRY(cudaMemsetAsync(d_count, 0, nfamilies * sizeof(int32_t), ctx->stream));
    const int tb = 256, nb = (nreads + tb - 1) / tb;
    fam_count_kernel<<<nb, tb, 0, ctx->stream>>>(d_read_family, nreads, nfamilies, d_count);
    fam_scan_kernel<<<1, 32, 0, ctx->stream>>>(d_count, nfamilies, d_fam_start, d_chunk_start, d_cursor);
    fam_scatter_kernel<<<nb, tb, 0, ctx->stream>>>(d_read_family, nreads, nfamilies, d_cursor, d_order, p.out);
    CUDA_TRY(cudaGetLastError());
    ctx->launches += 3;
    p.order = d_order; p.fam_start = d_fam_start; p.chunk_start = d_chunk_start;
    p.nfamilies = nfamilies; p.go = gap_open; p.ge = gap_extend; p.max_rows = max_rows;
    p.allow_fast = allow_fast ? 1 : 0;
    p.stats = d_stats;
    CUDA_TRY(cudaMemcpyToSymbolAsync(c_fmat25, mat25, 25, 0, cudaMemcpyHostToDevice, ctx->stream));
    const int nitems_bound = nreads / 32 + nfamilies + 1;
    int rc;
    ctx->mark(0);
    if (need_generic) { if ((rc = launch_classify<0>(ctx, p, nitems_bound, smem))) return rc; }
#define LAUNCH_P(PP) if (pmask & (1u << PP)) { if ((rc = launch_classify<PP>(ctx, p, nitems_bound, smem))) return rc; }
    LAUNCH_P(1) LAUNCH_P(2) LAUNCH_P(3) LAUNCH_P(4) LAUNCH_P(5) LAUNCH_P(6)
    LAUNCH_P(7) LAUNCH_P(8) LAUNCH_P(9) LAUNCH_P(10) LAUNCH_P(11) LAUNCH_P(12)
#undef LAUNCH_P
    ctx->mark(1);
    return TREDSW_OK;
}

extern "C" int tredsw_classify_reads(tredsw_ctx *ctx, const int8_t *rbuf, const int64_t *roff, int32_t nreads,
                                     const int32_t *read_family, const tredsw_family *families,
                                     int32_t nfamilies, const int8_t *mat25, int gap_open, int gap_extend,
                                     uint32_t flags, int32_t *out, int64_t *stats) {
    if (!ctx) { tredsw_set_error("null context"); return TREDSW_ERR_ARG; }
    if (nreads < 0 || nfamilies <= 0 || !families || !mat25 || !out) { tredsw_set_error("bad arguments"); return TREDSW_ERR_ARG; }
    if (nreads == 0) return TREDSW_OK;
    if (dev_ptrs(flags)) { tredsw_set_error("tredsw_classify_reads takes host buffers; device-resident batches go through tredsw_genotype_batch"); return TREDSW_ERR_UNSUPPORTED; }
    std::lock_guard<std::mutex> lock(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    int max_m = 0;
    for (int i = 0; i < nreads; ++i) { int l = (int)(roff[i + 1] - roff[i]); if (l > max_m) max_m = l; }
    int rc;
    const int8_t *d_rbuf; const int64_t *d_roff; const int32_t *d_rfam; const tredsw_family *d_fam;
    if ((rc = stage_in(ctx, ctx->d_q, rbuf, (size_t)roff[nreads], flags, &d_rbuf))) return rc;
    if ((rc = stage_in(ctx, ctx->d_qoff, roff, (size_t)nreads + 1, flags, &d_roff))) return rc;
    if ((rc = stage_in(ctx, ctx->d_rfam, read_family, (size_t)nreads, flags, &d_rfam))) return rc;
    if ((rc = stage_in(ctx, ctx->d_fam, families, (size_t)nfamilies, flags, &d_fam))) return rc;
    if ((rc = ctx->d_out.ensure((size_t)nreads * 8 * sizeof(int32_t)))) return rc;
    if ((rc = ctx->d_work.ensure(((size_t)4 * nfamilies + 2 + nreads) * sizeof(int32_t)))) return rc;
    if ((rc = ctx->d_stats.ensure(4 * sizeof(unsigned long long)))) return rc;
    CUDA_TRY(cudaMemsetAsync(ctx->d_stats.p, 0, 4 * sizeof(unsigned long long), ctx->stream));
    if ((rc = tredsw_internal_classify(ctx, d_rbuf, d_roff, nreads, d_rfam, d_fam, families, nfamilies, max_m, mat25,
                                       gap_open, gap_extend, ctx->d_work.as<int32_t>(), ctx->d_out.as<int32_t>(),
                                       stats ? ctx->d_stats.as<unsigned long long>() : nullptr))) return rc;
    CUDA_TRY(cudaMemcpyAsync(out, ctx->d_out.p, (size_t)nreads * 8 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (stats) CUDA_TRY(cudaMemcpyAsync(stats, ctx->d_stats.p, 4 * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return TREDSW_OK;
}

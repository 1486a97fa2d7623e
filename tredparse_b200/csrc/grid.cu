// grid.cu — the (h1, h2) log-likelihood surface of IntegratedCaller and its reductions, FP64.
//
// Replaces the double loop of tredparse/models.py:260-273 (evaluate_spanning / evaluate_partial /
// evaluate_rept / PEMaxLikModel.evaluate at every candidate pair) and the reductions of
// models.py:277-302, 304-317, 342-368 (max, arg-max with key (ml, -h1), exp-normalised marginals, the sparse
// joint posterior, PP sums).
//
// The reference materialises length-1000 probability vectors per candidate allele and takes the log of whole
// vectors at every grid point; only the entries at the observed keys are ever used.  Here a point is evaluated
// from the closed forms of SURVEY.md Appendix B:
//   PS(h)[k]  spanning pdf (models.py:149-168, quirks Q6/Q7)     PT(h)[k] partial pdf (:170-180, Q8)
//   alpha     mixing weights (:182-190)                           R(h)[x]  rolled PE pdf (:441-458)
// and a sum of logarithms  sum_k c_k log(max(v_k, eps))  is taken as the logarithm of the running PRODUCT of the
// v_k (flushed every 64 factors: v_k >= eps = e^-10 keeps 64 factors above e^-640) plus log(eps) times the
// number of floored factors — one FP64 log per term and point instead of one per observed key (rounding
// differs from the reference's sum by ~1e-16 relative; the parity bar is 1e-9).
//
// THE SURFACE IS NOT MATERIALISED (unless the caller asks for it).  What the caller of the reference gets are
// the call, the marginals P_h1 / P_h2, the sparse joint P_h1h2 and the PP sums; so:
//   * small surfaces (<= 2048 points, every haploid problem): one warp per problem evaluates, reduces and emits
//     in one kernel (grid_small_kernel), the ml values parked in the problem's scratch slot (L1/L2);
//   * large surfaces: row-structured.  A warp owns rows (one h1), its lanes walk the columns (h2) in chunks of
//     256; everything that depends on h1 only is hoisted out of the columns.  Where the longer allele lies
//     beyond every observed key, the partial clamp and the read length (column index >= fam, the MID region)
//     the spanning and partial terms depend on h1 only; where it is also shifted past the KDE support
//     (>= fa2, the FAR region) so does the paired-end term; the repeat-only term is a function of
//     dsum = max(h1-L,1) + max(h2-L,1) everywhere.  With per-row tables {c12(h1), pe(h1)} and rept[dsum] a far
//     point is  ml = (c12 + rept[dsum]) + pe  — on a --fullsearch / long-expansion grid > 90 % of the points —
//     and its weight exp(ml - max) = exp(c12 + pe - max) * exp(rept[dsum]): one multiply with a per-row factor
//     and a tabulated exp(rept) (rept lies in [-100, 0], so neither factor can overflow).
//     Pass A (grid_rows_eval_kernel) evaluates the near / mid points into the scratch slot and finds the
//     maximum; pass B (grid_rows_reduce_kernel) accumulates row sums, column sums (registers -> shared memory ->
//     per-CTA partials, summed by the last CTA of the problem in rank order: deterministic), the PP sums, the
//     arg-max (Q10) and emits the joint-posterior entries >= e^-10.
// DRAM traffic is the near-region scratch (a few % of the surface, L2-resident) plus the tables.
#include "internal.cuh"
#include "kde.cuh"
#include <math.h>

namespace {

constexpr int SPAN = 1000;
constexpr int NSTEP = 37;
constexpr int DEV = 18;
constexpr long long SMALL_LIMIT = 2048;   // <= : one warp per problem
constexpr int NC_MAX = 32;                // CTAs per large surface (rows are dealt round-robin in groups of 8)
constexpr int CHUNK = 256;                // columns per chunk: 8 per lane
constexpr int FILL_SPLIT = 8;             // blocks per large surface filling its tables

// per large surface: constants, thresholds, table layout, cross-CTA reduction state (zeroed every call)
struct BigInfo {
    unsigned long long maxkey;   // ordered key of the surface maximum (pass A, atomicMax); 0 = no point
    unsigned long long argkey;   // pass B: min over the points at the maximum of (h1 << 40 | row-major index)
    long long off;               // table arena offset (doubles)
    double lgamma_k1;            // lgamma(n_rept + 1)
    double sig_mp;               // sigma(max_partial)
    int tmin;                    // smallest pair length >= MINPE (after numpy's negative-index wrap); INT_MAX if none
    int fam, fa2;                // first column of the mid / far region
    int ok;                      // row / rept / sigma tables present
    int npe;                     // paired-end operand tables present (= n_target)
    int sorted;                  // both candidate lists non-decreasing: rows skip the h1 > h2 columns wholesale
    int nd;                      // entries of rept2[]
    int hrep;                    // a far allele (the largest h2)
    int mx1;                     // largest h1
    int nc;                      // CTAs working on this surface
    int nb1, nb2;                // length of the sorted base part of the h1 / h2 list (duplicates live beyond it)
    int npoints;                 // evaluated points (pass A)
    unsigned int done_b;         // CTAs of pass B that have finished
    int alloc_fail;              // the table arena was too small even for the mandatory part
    int pad_;
};

struct GridParams {
    const tredsw_grid_problem *prob;
    const int32_t *ipool;
    const double *dpool;
    double *surface;             // per-problem slots (off_surface): scratch, or the full surface when materialise
    double *marg;
    tredsw_grid_result *res;
    double small_value, log_small;
    int nproblems;
    int materialise;             // write every point (and -inf where h1 > h2) into the surface slots
    BigInfo *big;                // [nproblems]
    int *lists;                  // [0] number of large surfaces, [1 + i] their problem indices
    double *ftab;                // table arena
    long long ftab_cap;
    unsigned long long *fcursor; // [0] arena cursor, [1] overflow flag
    tredsw_posterior *post;      // sparse joint-posterior entries (optional)
    long long post_cap;
    unsigned long long *post_cursor;
};

// table layout of one large surface, in doubles from BigInfo.off
struct Tab {
    long long rows, sig1, sig2, rept2, dup, R1, R2, colpart, partial, end;
};
__host__ __device__ inline Tab tab_layout(int n1, int n2, int nd, int npe, int nc, bool with_tables) {
    Tab t;
    long long o = 0;
    t.colpart = o; o += (long long)nc * n2;
    t.partial = o; o += (long long)nc * 4;
    t.dup = o; o += ((long long)n1 + n2 + 1) / 2;            // int32 flags, two per double
    t.rows = o; if (with_tables) o += 3LL * n1;              // {c12, pe_far, f} per row
    t.sig1 = o; if (with_tables) o += n1;
    t.sig2 = o; if (with_tables) o += n2;
    o = (o + 1) & ~1LL;                                      // 16-byte aligned pairs
    t.rept2 = o; if (with_tables) o += 2LL * nd;             // {rept, exp(rept)} per dsum
    t.R1 = o; if (with_tables) o += (long long)npe * n1;
    t.R2 = o; if (with_tables) o += (long long)npe * n2;
    t.end = (o + 1) & ~1LL;
    return t;
}

// order-preserving map double -> u64 (for atomicMax); 0 is below every value
__device__ __forceinline__ unsigned long long ord_key(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}
__device__ __forceinline__ double ord_val(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffULL) : ~k;
    return __longlong_as_double((long long)b);
}

__device__ __forceinline__ double sigma_h(const tredsw_grid_problem &P, int h) {
    double z = __dadd_rn(P.stutter_a, __dmul_rn(P.stutter_w2, (double)(h / P.period)));
    z = __dadd_rn(__dadd_rn(z, P.stutter_c3), P.stutter_c4);
    return 1.0 / (1.0 + exp(-1.0 * z));
}

// spanning pdf of allele h at key k given sig = sigma(h)
__device__ __forceinline__ double pdf_span(const double *step, int h, double sig, int k) {
    if (k < 0 || k >= SPAN) return 0.0;
    int idx;
    if (h + DEV + 1 <= SPAN) {
        idx = k - h + DEV;                      // also covers the low-end clip (h < 18): tail of p
    } else {
        if (k < h - DEV) return 0.0;            // high-end clip copies the *tail* of p as well (Q6)
        idx = k - (SPAN - NSTEP);
    }
    if (idx < 0 || idx >= NSTEP) return 0.0;
    return idx == DEV ? (1.0 - sig) : __dmul_rn(step[idx], sig);
}

__device__ __forceinline__ double pdf_part(const double *step, int hc, double sig_c, double c, int k) {
    const double v = (k < hc) ? c : 0.0;
    return __dadd_rn(v, __dmul_rn(c, pdf_span(step, hc, sig_c, k)));
}

__device__ __forceinline__ double pe_roll(const double *pdf, int h, int ref, int minpe, int x, double eps) {
    if (x < minpe) return eps;
    const int y = x + h - ref;
    if (y < 0 || y >= SPAN) return eps;
    return pdf[y];
}

// sum_k c_k log(v_k) as the log of a running product; every factor lies in [eps, 1], so 64 of them stay above
// e^-640.  The explicit intrinsics keep the compiler from contracting differently in different call sites: the
// table-driven and the direct evaluation of a point have to give the same bits.
struct LogProd {
    double acc, prod;
    int room;
    __device__ __forceinline__ LogProd() : acc(0.0), prod(1.0), room(64) {}
    __device__ __forceinline__ void flush() {
        if (room != 64) { acc = __dadd_rn(acc, log(prod)); prod = 1.0; room = 64; }
    }
    __device__ __forceinline__ void mul(double v) {
        prod = __dmul_rn(prod, v);
        if (--room == 0) flush();
    }
    __device__ __forceinline__ void mulc(double v, int c) {
        if (c > 4) acc = __dadd_rn(acc, __dmul_rn((double)c, log(v)));
        else for (int q = 0; q < c; ++q) mul(v);
    }
    __device__ __forceinline__ double done() { flush(); return acc; }
};

// ---- the four terms of one point -------------------------------------------------------------------------
// sg1 / sg2: sigma(h1) / sigma(h2) when the caller has them (row / column tables), negative = compute on
// first use.  A key farther than 18 from both alleles sees 0 -> floored to eps.
__device__ __forceinline__ double term_span(const tredsw_grid_problem &P, const GridParams &g, int h1, int h2,
                                            double sg1, double sg2) {
    if (P.n_span <= 0) return 0.0;
    const int32_t *skey = g.ipool + P.off_span, *scnt = skey + P.n_span;
    const double *step = g.dpool + P.off_step;
    const double eps = g.small_value;
    const int t2 = P.readlen - 18;
    const int s1 = max(0, t2 - h1), s2 = max(0, t2 - h2);
    const double a = (s1 + s2) ? (double)s1 * 1.0 / (double)(s1 + s2) : 0.5;
    const double b = 1.0 - a;
    const int lo = min(h1, h2) - DEV, hi = max(h1, h2) + DEV;
    LogProd lp;
    int nfloor = 0;
    for (int i = 0; i < P.n_span; ++i) {
        const int k = skey[i], c = scnt[i];
        if (k < lo || k > hi) { nfloor += c; continue; }
        if (sg1 < 0.0) sg1 = sigma_h(P, h1);
        if (sg2 < 0.0) sg2 = sigma_h(P, h2);
        const double v = __dadd_rn(__dmul_rn(a, pdf_span(step, h1, sg1, k)), __dmul_rn(b, pdf_span(step, h2, sg2, k)));
        if (v <= eps) { nfloor += c; continue; }
        lp.mulc(v, c);
    }
    return __dadd_rn(lp.done(), __dmul_rn((double)nfloor, g.log_small));
}

// sgc1 / sgc2: sigma of the alleles clamped to max_partial (negative = compute on first use).  A partial key
// below both clamped alleles by more than 18 sees alpha*c1 + (1-alpha)*c2, one above both sees 0.
__device__ __forceinline__ double term_part(const tredsw_grid_problem &P, const GridParams &g, int h1, int h2,
                                            double sgc1, double sgc2, double sig_mp) {
    if (P.n_part <= 0) return 0.0;
    const int32_t *pkey = g.ipool + P.off_part, *pcnt = pkey + P.n_part;
    const double *step = g.dpool + P.off_step;
    const double eps = g.small_value;
    const int t1 = P.readlen - 9;
    const int s1 = min(h1, t1), s2 = min(h2, t1);
    const double a = (s1 + s2) ? (double)s1 * 1.0 / (double)(s1 + s2) : 0.5;
    const double b = 1.0 - a;
    const int hc1 = min(h1, P.max_partial), hc2 = min(h2, P.max_partial);
    const double c1 = 1.0 / (double)(hc1 + 1), c2 = 1.0 / (double)(hc2 + 1);
    const int lo = min(hc1, hc2) - DEV, hi = max(hc1, hc2) + DEV;
    const double v_bulk = __dadd_rn(__dmul_rn(a, c1), __dmul_rn(b, c2));   // p1 = c1 + c1 * 0, p2 = c2 + c2 * 0
    LogProd lp;
    int nfloor = 0, nbulk = 0;
    for (int i = 0; i < P.n_part; ++i) {
        const int k = pkey[i], c = pcnt[i];
        if (k < lo) { nbulk += c; continue; }
        if (k > hi) { nfloor += c; continue; }
        if (sgc1 < 0.0) sgc1 = hc1 == P.max_partial ? sig_mp : sigma_h(P, hc1);
        if (sgc2 < 0.0) sgc2 = hc2 == P.max_partial ? sig_mp : sigma_h(P, hc2);
        const double v = __dadd_rn(__dmul_rn(a, pdf_part(step, hc1, sgc1, c1, k)), __dmul_rn(b, pdf_part(step, hc2, sgc2, c2, k)));
        if (v <= eps) { nfloor += c; continue; }
        lp.mulc(v, c);
    }
    double acc = lp.done();
    if (nbulk) {
        if (v_bulk <= eps) nfloor += nbulk;
        else acc = __dadd_rn(acc, __dmul_rn((double)nbulk, log(v_bulk)));
    }
    return __dadd_rn(acc, __dmul_rn((double)nfloor, g.log_small));
}

// repeat-only reads: Poisson (scipy: exp(xlogy(k, mu) - gammaln(k + 1) - mu)); dsum = max(h1-L,1) + max(h2-L,1)
__device__ __forceinline__ double term_rept(const tredsw_grid_problem &P, int dsum, double lgamma_k1) {
    const double mu = (double)dsum * P.half_depth / (double)P.readlen;
    const double kk = (double)P.n_rept;
    const double xl = (P.n_rept == 0) ? 0.0 : __dmul_rn(kk, log(mu));
    const double pk = __dsub_rn(__dsub_rn(xl, lgamma_k1), mu);
    // log(max(exp(pk), e^-100)): log(exp(pk)) is pk to within an ulp of the pmf (~1e-16 absolute on a term
    // of magnitude 0.1..100, far inside the 1e-9 relative bar)
    return pk > -100.0 ? pk : -100.0;
}

__device__ __forceinline__ double term_pe(const tredsw_grid_problem &P, const GridParams &g, int h1, int h2, int tmin) {
    if (!P.run_pe) return 0.0;
    const double eps = g.small_value;
    const long long off1 = (long long)h1 - P.pe_ref, off2 = (long long)h2 - P.pe_ref;
    // every pair length is below MINPE or shifted past the end of the support: 0.5*eps + 0.5*eps = eps
    if (tmin == 0x7fffffff || (tmin + off1 >= SPAN && tmin + off2 >= SPAN)) return __dmul_rn((double)P.n_target, g.log_small);
    const double *pdf = g.dpool + P.off_pdf;
    const int32_t *tl = g.ipool + P.off_target;
    LogProd lp;
    int nfloor = 0;
    for (int i = 0; i < P.n_target; ++i) {
        int x = tl[i];
        if (x < 0) x += SPAN;                   // numpy negative-index wrap (models.py:473)
        const double r1 = pe_roll(pdf, h1, P.pe_ref, P.pe_minpe, x, eps);
        const double r2 = pe_roll(pdf, h2, P.pe_ref, P.pe_minpe, x, eps);
        const double v = __dadd_rn(__dmul_rn(0.5, r1), __dmul_rn(0.5, r2));
        if (v <= eps) ++nfloor; else lp.mul(v);
    }
    return __dadd_rn(lp.done(), __dmul_rn((double)nfloor, g.log_small));
}

// term_pe from the tabulated operands R1[t][i1] / R2[t][i2] (exactly what pe_roll returns)
__device__ __forceinline__ double term_pe_tab(const double *R1, const double *R2, int n1, int n2, int npe, int i1, int i2,
                                              double eps, double log_small) {
    LogProd lp;
    int nfloor = 0;
    R1 += i1; R2 += i2;
    for (int t = 0; t < npe; ++t) {
        const double v = __dadd_rn(__dmul_rn(0.5, R1[(long long)t * n1]), __dmul_rn(0.5, R2[(long long)t * n2]));
        if (v <= eps) ++nfloor; else lp.mul(v);
    }
    return __dadd_rn(lp.done(), __dmul_rn((double)nfloor, log_small));
}

struct ProblemConsts {
    double lgamma_k1;      // lgamma(n_rept + 1) of the Poisson term
    double sig_mp;         // sigma(max_partial): the stutter probability of every allele clamped to max_partial
    int tmin;              // smallest pair length >= MINPE (after numpy's negative-index wrap); INT_MAX if none
};

__device__ __forceinline__ double point_ml(const tredsw_grid_problem &P, const GridParams &g, int h1, int h2,
                                           const ProblemConsts &T, double sg1, double sg2, double sgc1, double sgc2) {
    double ml = term_span(P, g, h1, h2, sg1, sg2);
    ml = __dadd_rn(ml, term_part(P, g, h1, h2, sgc1, sgc2, T.sig_mp));
    ml = __dadd_rn(ml, term_rept(P, max(h1 - P.readlen, 1) + max(h2 - P.readlen, 1), T.lgamma_k1));
    ml = __dadd_rn(ml, term_pe(P, g, h1, h2, T.tmin));
    return ml;
}

// (u1, u2 = h1 / period, h2 / period: division is monotone, so min / max commute with it)
__device__ __forceinline__ bool pathological_u(const tredsw_grid_problem &P, int u1, int u2) {
    const int lo = min(u1, u2), hi = max(u1, u2);
    if (P.expansion) return P.recessive ? (lo >= P.cutoff_risk) : (hi >= P.cutoff_risk);
    return P.recessive ? (hi <= P.cutoff_risk) : (lo <= P.cutoff_risk);
}

// A candidate list is a sorted base part [0, nb) followed by an ascending extension (models.py:250-257; with
// --fullsearch the whole list is ascending and nb = n).  An extension entry repeating a base value is the
// second occurrence of that allele (Q9): its points are evaluated and counted again in the marginals and the
// PP sums, but the joint posterior is a dict keyed by (h1, h2) and holds them once.
__device__ __forceinline__ int base_len(const int32_t *hs, int n, int lane) {       // warp-cooperative
    int nb = n;
    for (int i = 1 + lane; i < n; i += 32) if (hs[i] <= hs[i - 1]) { nb = min(nb, i); break; }
    return __reduce_min_sync(0xffffffffu, nb);
}
__device__ __forceinline__ bool is_second_occurrence(const int32_t *hs, int nb, int i) {
    if (i < nb) return false;
    const int v = hs[i];
    if (nb == 0 || v > hs[nb - 1]) return false;
    for (int j = 0; j < nb; ++j) if (hs[j] == v) return true;
    return false;
}

__device__ __forceinline__ void emit_joint(const GridParams &g, int pi, int u1, int u2, double w) {
    if (!g.post) return;
    const unsigned long long at = atomicAdd(g.post_cursor, 1ULL);
    if ((long long)at < g.post_cap) {
        tredsw_posterior e;
        e.problem = pi; e.kind = TREDSW_POST_JOINT; e.a = u1; e.b = u2; e.p = w;
        g.post[at] = e;
    }
}

// ==========================================================================================================
// small surfaces: one warp per problem — evaluate, reduce, emit.  Large ones are registered for the row kernels.
// ==========================================================================================================
__global__ void __launch_bounds__(256) grid_small_kernel(GridParams g) {
    const int lane = threadIdx.x & 31;
    const int pi = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (pi >= g.nproblems) return;
    const tredsw_grid_problem &P = g.prob[pi];
    const int n1 = P.n_h1, n2 = P.n_h2;
    const long long total = n2 > 0 ? (long long)n1 * n2 : 0;
    const int32_t *h1s = g.ipool + P.off_h1, *h2s = g.ipool + P.off_h2;
    const bool haploid = P.ploidy == 1;
    if (!haploid && total > SMALL_LIMIT) {
        // ---- register a large surface: list slot, list maxima, table allocation --------------------------
        int mx1 = -0x7fffffff, mx2 = -0x7fffffff;
        for (int i = lane; i < n1; i += 32) mx1 = max(mx1, h1s[i]);
        for (int i = lane; i < n2; i += 32) mx2 = max(mx2, h2s[i]);
        mx1 = __reduce_max_sync(0xffffffffu, mx1); mx2 = __reduce_max_sync(0xffffffffu, mx2);
        const int nb1 = base_len(h1s, n1, lane), nb2 = base_len(h2s, n2, lane);
        if (lane == 0) {
            BigInfo B;
            memset(&B, 0, sizeof(B));
            B.argkey = ~0ULL;
            B.mx1 = mx1; B.hrep = mx2; B.nb1 = nb1; B.nb2 = nb2;
            B.nd = max(mx1 - P.readlen, 1) + max(mx2 - P.readlen, 1) - 1;
            B.nc = min(NC_MAX, max(1, (n1 + 31) / 32));
            const int npe = (P.run_pe && P.n_target > 0) ? P.n_target : 0;
            const Tab full = tab_layout(n1, n2, B.nd, npe, B.nc, true), nope = tab_layout(n1, n2, B.nd, 0, B.nc, true),
                      bare = tab_layout(n1, n2, B.nd, 0, B.nc, false);
            // the reduction scratch is mandatory; the tables are taken when the arena has room for them
            unsigned long long off = atomicAdd(&g.fcursor[0], (unsigned long long)full.end);
            if ((long long)(off + full.end) <= g.ftab_cap) { B.ok = 1; B.npe = npe; }
            else {
                off = atomicAdd(&g.fcursor[0], (unsigned long long)nope.end);
                if ((long long)(off + nope.end) <= g.ftab_cap) { B.ok = 1; B.npe = 0; }
                else {
                    off = atomicAdd(&g.fcursor[0], (unsigned long long)bare.end);
                    if ((long long)(off + bare.end) > g.ftab_cap) { B.alloc_fail = 1; atomicExch(&g.fcursor[1], 1ULL); off = 0; }
                }
            }
            B.off = (long long)off;
            g.big[pi] = B;
            if (!B.alloc_fail) { const int slot = atomicAdd(&g.lists[0], 1); g.lists[1 + slot] = pi; }
            else {
                tredsw_grid_result r;
                memset(&r, 0, sizeof(r));
                r.arg_i1 = r.arg_i2 = -1; r.n_points = -1;
                g.res[pi] = r;
            }
        }
        return;
    }
    // ---- evaluate --------------------------------------------------------------------------------------
    ProblemConsts T;
    T.lgamma_k1 = lgamma((double)P.n_rept + 1.0);
    T.sig_mp = sigma_h(P, P.max_partial);
    int tmin = 0x7fffffff;
    if (P.run_pe) {
        const int32_t *tl = g.ipool + P.off_target;
        for (int i = lane; i < P.n_target; i += 32) { int x = tl[i]; if (x < 0) x += SPAN; if (x >= P.pe_minpe && x < tmin) tmin = x; }
        tmin = __reduce_min_sync(0xffffffffu, tmin);
    }
    T.tmin = tmin;
    double *surf = g.surface + P.off_surface;
    double best_ml = -INFINITY;
    int best_h1 = 0x7fffffff, best_t = 0x7fffffff, cnt = 0;
    for (long long tt = lane; tt < total; tt += 32) {
        const int t = (int)tt;
        const int i1 = t / n2, i2 = t - i1 * n2;
        const int h1 = h1s[i1];
        const int h2 = haploid ? h1 : h2s[i2];
        double ml = -INFINITY;
        if (h1 <= h2) {
            ml = point_ml(P, g, h1, h2, T, -1.0, -1.0, -1.0, -1.0);
            ++cnt;
            if (ml > best_ml || (ml == best_ml && (h1 < best_h1 || (h1 == best_h1 && t < best_t)))) { best_ml = ml; best_h1 = h1; best_t = t; }
        }
        surf[t] = ml;
    }
    for (int d = 16; d > 0; d >>= 1) {
        const double oml = __shfl_xor_sync(0xffffffffu, best_ml, d);
        const int oh = __shfl_xor_sync(0xffffffffu, best_h1, d), ot = __shfl_xor_sync(0xffffffffu, best_t, d);
        if (oml > best_ml || (oml == best_ml && (oh < best_h1 || (oh == best_h1 && ot < best_t)))) { best_ml = oml; best_h1 = oh; best_t = ot; }
        cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    }
    __syncwarp();
    // ---- reduce ----------------------------------------------------------------------------------------
    double *ph1 = g.marg + P.off_ph1, *ph2 = g.marg + P.off_ph2;
    for (int i2 = lane; i2 < n2; i2 += 32) ph2[i2] = 0.0;
    const int nb1 = cnt ? base_len(h1s, n1, lane) : n1;
    const int nb2 = (cnt && !haploid) ? base_len(h2s, n2, lane) : n2;
    double sum_all = 0.0, sum_path = 0.0, sum_uniq = 0.0;
    const double eps = g.small_value;
    for (int i1 = 0; i1 < n1; ++i1) {
        const int h1 = h1s[i1], u1 = h1 / P.period;
        const bool dup1 = is_second_occurrence(h1s, nb1, i1);
        double acc = 0.0, accp = 0.0, accu = 0.0;
        for (int i2 = lane; i2 < n2; i2 += 32) {           // column i2 always belongs to lane i2 % 32: the plain
            const double ml = surf[i1 * n2 + i2];          // read-modify-write of ph2 below is race-free and ordered
            if (ml == -INFINITY) continue;
            const double w = exp(ml - best_ml);
            const int u2 = haploid ? u1 : h2s[i2] / P.period;
            acc += w; ph2[i2] += w;
            if (pathological_u(P, u1, u2)) accp += w;
            if (!(dup1 || (!haploid && is_second_occurrence(h2s, nb2, i2)))) {
                accu += w;
                if (w >= eps) emit_joint(g, pi, u1, u2, w);
            }
        }
        for (int d = 16; d > 0; d >>= 1) {
            acc += __shfl_down_sync(0xffffffffu, acc, d);
            accp += __shfl_down_sync(0xffffffffu, accp, d);
            accu += __shfl_down_sync(0xffffffffu, accu, d);
        }
        if (lane == 0) { ph1[i1] = acc; sum_all += acc; sum_path += accp; sum_uniq += accu; }
    }
    if (lane == 0) {
        tredsw_grid_result r;
        r.max_ml = best_ml; r.sum_all = sum_all; r.sum_path = sum_path; r.sum_uniq = sum_uniq;
        r.arg_i1 = cnt ? best_t / n2 : -1;
        r.arg_i2 = cnt ? best_t % n2 : -1;
        r.n_points = cnt; r.pad = 0;
        g.res[pi] = r;
    }
}

// ==========================================================================================================
// large surfaces
// ==========================================================================================================
// block-wide max of an int (256 threads); every thread gets the result
__device__ __forceinline__ int block_max_int(int v, int *s8) {
    v = __reduce_max_sync(0xffffffffu, v);
    __syncthreads();                                   // s8 may still be read from the previous call
    if ((threadIdx.x & 31) == 0) s8[threadIdx.x >> 5] = v;
    __syncthreads();
    int r = s8[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) r = max(r, s8[w]);
    return r;
}

// FILL_SPLIT blocks per large surface: each derives the thresholds (cheap, block-parallel scans — no block waits
// for another), part 0 publishes them, all fill their share of the tables.
__global__ void __launch_bounds__(256) grid_big_setup_kernel(GridParams g) {
    __shared__ int s8[8];
    const int nbig = g.lists[0];
    const int tid = threadIdx.x;
    for (int kk = blockIdx.x; kk < nbig * FILL_SPLIT; kk += gridDim.x) {
        const int pi = g.lists[1 + kk / FILL_SPLIT], part = kk % FILL_SPLIT;
        const tredsw_grid_problem &P = g.prob[pi];
        BigInfo &B = g.big[pi];
        const int n1 = P.n_h1, n2 = P.n_h2;
        const int32_t *h1s = g.ipool + P.off_h1, *h2s = g.ipool + P.off_h2;
        const int32_t *skey = g.ipool + P.off_span, *tl = g.ipool + P.off_target;
        int ks = -1000000, ntmin = -0x7fffffff, unsorted = 0;
        for (int i = tid; i < P.n_span; i += 256) ks = max(ks, skey[i]);
        if (P.run_pe)
            for (int i = tid; i < P.n_target; i += 256) { int x = tl[i]; if (x < 0) x += SPAN; if (x >= P.pe_minpe) ntmin = max(ntmin, -x); }
        for (int i = tid + 1; i < n1; i += 256) if (h1s[i - 1] > h1s[i]) unsorted = 1;
        for (int i = tid + 1; i < n2; i += 256) if (h2s[i - 1] > h2s[i]) unsorted = 1;
        ks = block_max_int(ks, s8); ntmin = block_max_int(ntmin, s8); unsorted = block_max_int(unsorted, s8);
        const int tmin = ntmin == -0x7fffffff ? 0x7fffffff : -ntmin;
        const int t1 = P.readlen - 9;
        // h2 >= H1: no spanning key within 18, partial clamp and mixing weights saturated, h2 > readlen
        const int H1 = max(max(ks + DEV + 1, P.max_partial), max(t1, P.readlen + 1));
        // h2 >= H2: every pair length >= MINPE is shifted past the support as well
        long long H2 = H1;
        if (P.run_pe && tmin != 0x7fffffff) H2 = max((long long)H1, (long long)P.pe_ref + SPAN - tmin);
        int l1 = -1, l2 = -1;                          // last column below H1 / H2
        for (int i = tid; i < n2; i += 256) { const int h = h2s[i]; if (h < H1) l1 = max(l1, i); if (h < H2) l2 = max(l2, i); }
        l1 = block_max_int(l1, s8); l2 = block_max_int(l2, s8);
        const int fam = B.ok ? l1 + 1 : n2, fa2 = B.ok ? l2 + 1 : n2;
        const double lgk = lgamma((double)P.n_rept + 1.0), sig_mp = sigma_h(P, P.max_partial);
        if (part == 0 && tid == 0) {
            B.lgamma_k1 = lgk; B.sig_mp = sig_mp; B.tmin = tmin; B.fam = fam; B.fa2 = fa2; B.sorted = unsorted ? 0 : 1;
        }
        const Tab tb = tab_layout(n1, n2, B.nd, B.npe, B.nc, B.ok != 0);
        double *tab = g.ftab + B.off;
        int32_t *dup = reinterpret_cast<int32_t *>(tab + tb.dup);
        for (int e = part * 256 + tid; e < n1 + n2; e += 256 * FILL_SPLIT)
            dup[e] = e < n1 ? (int)is_second_occurrence(h1s, B.nb1, e) : (int)is_second_occurrence(h2s, B.nb2, e - n1);
        if (!B.ok) continue;
        const int hrep = B.hrep, npe = B.npe, nd = B.nd;
        const double *pdf = g.dpool + P.off_pdf;
        const long long o_s1 = 0, o_s2 = o_s1 + n1, o_rows = o_s2 + n2, o_rept = o_rows + n1, o_R1 = o_rept + nd,
                        o_R2 = o_R1 + (long long)npe * n1, n_all = o_R2 + (long long)npe * fa2;
        for (long long e = part * 256 + tid; e < n_all; e += 256 * FILL_SPLIT) {
            if (e < o_s2) tab[tb.sig1 + e] = sigma_h(P, h1s[e]);
            else if (e < o_rows) tab[tb.sig2 + (e - o_s2)] = sigma_h(P, h2s[e - o_s2]);
            else if (e < o_rept) {
                const int i1 = (int)(e - o_rows), h1 = h1s[i1];
                const int hc1 = min(h1, P.max_partial);
                const double s1 = sigma_h(P, h1), sc1 = hc1 == P.max_partial ? sig_mp : s1;
                // (for a far column the second allele's sigma is never used: every key is farther than 18 from it;
                //  sig_mp stands in for the clamped one)
                tab[tb.rows + 3LL * i1] = __dadd_rn(term_span(P, g, h1, hrep, s1, 0.5), term_part(P, g, h1, hrep, sc1, sig_mp, sig_mp));
                tab[tb.rows + 3LL * i1 + 1] = term_pe(P, g, h1, hrep, tmin);
                tab[tb.rows + 3LL * i1 + 2] = 0.0;
            } else if (e < o_R1) {
                const int d = (int)(e - o_rept);
                const double r = term_rept(P, d + 2, lgk);
                tab[tb.rept2 + 2LL * d] = r;
                tab[tb.rept2 + 2LL * d + 1] = exp(r);
            } else {
                // operands of the two-dimensional paired-end term, exactly as term_pe obtains them
                long long idx = e - o_R1;
                const bool first = e < o_R2;
                if (!first) idx = e - o_R2;
                const int width = first ? n1 : fa2;
                const int t = (int)(idx / width), i = (int)(idx - (long long)t * width);
                int x = tl[t];
                if (x < 0) x += SPAN;
                tab[(first ? tb.R1 : tb.R2) + idx] = pe_roll(pdf, first ? h1s[i] : h2s[i], P.pe_ref, P.pe_minpe, x, g.small_value);
            }
        }
    }
}

struct RowTables {
    const double *rows, *sig1, *sig2, *rept2, *R1, *R2;
    const int32_t *dup;
    double *colpart, *partial;
};
__device__ __forceinline__ RowTables row_tables(const GridParams &g, const BigInfo &B, int n1, int n2) {
    const Tab tb = tab_layout(n1, n2, B.nd, B.npe, B.nc, B.ok != 0);
    double *tab = g.ftab + B.off;
    RowTables r;
    r.rows = tab + tb.rows; r.sig1 = tab + tb.sig1; r.sig2 = tab + tb.sig2; r.rept2 = tab + tb.rept2;
    r.R1 = tab + tb.R1; r.R2 = tab + tb.R2; r.dup = reinterpret_cast<const int32_t *>(tab + tb.dup);
    r.colpart = tab + tb.colpart; r.partial = tab + tb.partial;
    return r;
}

// Pass A: near / mid points into the scratch slot, the maximum and the point count of every large surface.
// Work item = (surface, CTA rank c < nc): the row groups g = c, c + nc, ... of 8 rows (one per warp) — near rows
// (expensive) and far rows (cheap) are dealt evenly.
__global__ void __launch_bounds__(256) grid_rows_eval_kernel(GridParams g) {
    __shared__ unsigned long long s_key[8];
    __shared__ int s_cnt[8];
    const int nbig = g.lists[0];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int item = blockIdx.x; item < nbig * NC_MAX; item += gridDim.x) {
        const int pi = g.lists[1 + item / NC_MAX], c = item % NC_MAX;
        BigInfo &B = g.big[pi];
        const int nc = B.nc;
        if (c >= nc) continue;
        const tredsw_grid_problem &P = g.prob[pi];
        const int n1 = P.n_h1, n2 = P.n_h2, L = P.readlen;
        const int32_t *h1s = g.ipool + P.off_h1, *h2s = g.ipool + P.off_h2;
        double *surf = g.surface + P.off_surface, *ph1 = g.marg + P.off_ph1;
        const RowTables R = row_tables(g, B, n1, n2);
        const int ok = B.ok, fam = B.fam, fa2 = B.fa2, npe = B.npe, sorted = B.sorted;
        const bool mat = g.materialise != 0;
        ProblemConsts T;
        T.lgamma_k1 = B.lgamma_k1; T.sig_mp = B.sig_mp; T.tmin = B.tmin;
        const double eps = g.small_value, log_small = g.log_small;
        double mx = -INFINITY;
        int cnt = 0;
        for (int grp = c; grp * 8 < n1; grp += nc) {
            const int i1 = grp * 8 + warp;
            if (i1 < n1 && lane == 0) ph1[i1] = 0.0;
        }
        for (int cb = 0; cb < n2; cb += CHUNK) {
            const int cend = min(cb + CHUNK, n2);
            const int h2_last = sorted ? h2s[cend - 1] : 0x7fffffff;
            const bool all_far = ok && cb >= fa2;
            for (int grp = c; grp * 8 < n1; grp += nc) {
                const int i1 = grp * 8 + warp;
                if (i1 >= n1) continue;
                const int h1 = h1s[i1];
                const long long row = (long long)i1 * n2;
                if (h2_last < h1) {                              // the whole chunk has h1 > h2: not evaluated
                    if (mat) for (int col = cb + lane; col < cend; col += 32) surf[row + col] = -INFINITY;
                    continue;
                }
                const int dh1 = max(h1 - L, 1);
                double c12 = 0.0, pe_far = 0.0, sg1 = -1.0, sgc1 = -1.0;
                if (ok) {
                    c12 = R.rows[3LL * i1]; pe_far = R.rows[3LL * i1 + 1]; sg1 = R.sig1[i1];
                    sgc1 = min(h1, P.max_partial) == P.max_partial ? T.sig_mp : sg1;
                }
                if (all_far) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int col = cb + lane + 32 * j;
                        if (col >= cend) continue;
                        const int h2 = h2s[col];
                        double ml = -INFINITY;
                        if (h1 <= h2) {
                            ml = __dadd_rn(__dadd_rn(c12, R.rept2[2LL * (dh1 + max(h2 - L, 1) - 2)]), pe_far);
                            mx = fmax(mx, ml); ++cnt;
                        }
                        if (mat) surf[row + col] = ml;
                    }
                    continue;
                }
#pragma unroll 1
                for (int col = cb + lane; col < cend; col += 32) {
                    const int h2 = h2s[col];
                    double ml = -INFINITY;
                    if (h1 <= h2) {
                        if (ok) {
                            const double rp = R.rept2[2LL * (dh1 + max(h2 - L, 1) - 2)];
                            if (col >= fa2) ml = __dadd_rn(__dadd_rn(c12, rp), pe_far);
                            else {
                                const double pe = npe > 0 ? term_pe_tab(R.R1, R.R2, n1, fa2, npe, i1, col, eps, log_small)
                                                          : term_pe(P, g, h1, h2, T.tmin);
                                if (col >= fam) ml = __dadd_rn(__dadd_rn(c12, rp), pe);
                                else {
                                    const double sg2 = R.sig2[col];
                                    const double sgc2 = min(h2, P.max_partial) == P.max_partial ? T.sig_mp : sg2;
                                    ml = __dadd_rn(term_span(P, g, h1, h2, sg1, sg2), term_part(P, g, h1, h2, sgc1, sgc2, T.sig_mp));
                                    ml = __dadd_rn(__dadd_rn(ml, rp), pe);
                                }
                            }
                        } else {
                            ml = point_ml(P, g, h1, h2, T, -1.0, -1.0, -1.0, -1.0);
                        }
                        mx = fmax(mx, ml); ++cnt;
                        if (!(ok && col >= fa2) || mat) surf[row + col] = ml;
                    } else if (mat) surf[row + col] = ml;
                }
            }
        }
        unsigned long long key = cnt ? ord_key(mx) : 0ULL;
        for (int d = 16; d > 0; d >>= 1) {
            key = max(key, __shfl_xor_sync(0xffffffffu, key, d));
            cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
        }
        __syncthreads();
        if (lane == 0) { s_key[warp] = key; s_cnt[warp] = cnt; }
        __syncthreads();
        if (tid == 0) {
            unsigned long long m = s_key[0];
            int n = s_cnt[0];
#pragma unroll
            for (int w = 1; w < 8; ++w) { m = max(m, s_key[w]); n += s_cnt[w]; }
            if (m) atomicMax(&B.maxkey, m);
            if (n) atomicAdd(&B.npoints, n);
        }
    }
}

// Pass B: weights exp(ml - max) of every point -> row sums (P_h1), column sums (P_h2), PP sums, arg-max, joint
// entries.  Column sums: registers (8 columns per lane over the warp's rows) -> shared memory (8 warps, in warp
// order) -> this CTA's slice of colpart; the last CTA of a surface to finish adds the slices in rank order.
__global__ void __launch_bounds__(256) grid_rows_reduce_kernel(GridParams g) {
    __shared__ double s_col[8][CHUNK];
    __shared__ double s_sum[8][3];
    __shared__ unsigned long long s_arg[8];
    __shared__ int s_last;
    const int nbig = g.lists[0];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int item = blockIdx.x; item < nbig * NC_MAX; item += gridDim.x) {
        const int pi = g.lists[1 + item / NC_MAX], c = item % NC_MAX;
        BigInfo &B = g.big[pi];
        const int nc = B.nc;
        if (c >= nc) continue;
        const tredsw_grid_problem &P = g.prob[pi];
        const int n1 = P.n_h1, n2 = P.n_h2, L = P.readlen, K = P.period;
        const int32_t *h1s = g.ipool + P.off_h1, *h2s = g.ipool + P.off_h2;
        const double *surf = g.surface + P.off_surface;
        double *ph1 = g.marg + P.off_ph1, *ph2 = g.marg + P.off_ph2;
        const RowTables R = row_tables(g, B, n1, n2);
        double *rows_w = const_cast<double *>(R.rows);
        const int ok = B.ok, fa2 = B.fa2, sorted = B.sorted;
        const double eps = g.small_value;
        const unsigned long long maxkey = B.maxkey;
        const double M = maxkey ? ord_val(maxkey) : INFINITY;          // no point at all: nothing compares equal
        // per-row factor of the far weights
        if (ok)
            for (int grp = c; grp * 8 < n1; grp += nc) {
                const int i1 = grp * 8 + warp;
                if (i1 < n1 && lane == 0) rows_w[3LL * i1 + 2] = exp(__dadd_rn(R.rows[3LL * i1], R.rows[3LL * i1 + 1]) - M);
            }
        __syncwarp();
        double sum_all = 0.0, sum_path = 0.0, sum_uniq = 0.0;          // lane 0 of every warp
        unsigned long long argkey = ~0ULL;
        for (int cb = 0; cb < n2; cb += CHUNK) {
            const int cend = min(cb + CHUNK, n2);
            const int h2_last = sorted ? h2s[cend - 1] : 0x7fffffff;
            int h2v[8], dup2 = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int col = cb + lane + 32 * j;
                h2v[j] = col < cend ? h2s[col] : -0x7fffffff;
                if (col < cend && R.dup[n1 + col]) dup2 |= 1 << j;
            }
            double colacc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) colacc[j] = 0.0;
            for (int grp = c; grp * 8 < n1; grp += nc) {
                const int i1 = grp * 8 + warp;
                if (i1 >= n1) continue;
                const int h1 = h1s[i1];
                if (h2_last < h1) continue;
                const int u1 = h1 / K, dh1 = max(h1 - L, 1);
                const bool dup1 = R.dup[i1] != 0;
                const long long row = (long long)i1 * n2;
                double c12 = 0.0, pe_far = 0.0, f = 0.0;
                if (ok) { c12 = R.rows[3LL * i1]; pe_far = R.rows[3LL * i1 + 1]; f = R.rows[3LL * i1 + 2]; }
                double racc = 0.0, raccp = 0.0, raccu = 0.0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int h2 = h2v[j];
                    if (h1 > h2) continue;                                // not evaluated, or past the chunk
                    const int col = cb + lane + 32 * j;
                    double ml, w;
                    if (ok && col >= fa2) {
                        const double2 r = *reinterpret_cast<const double2 *>(R.rept2 + 2LL * (dh1 + max(h2 - L, 1) - 2));
                        ml = __dadd_rn(__dadd_rn(c12, r.x), pe_far);
                        w = f * r.y;
                    } else {
                        ml = surf[row + col];
                        const double d = ml - M;
                        w = d < -746.0 ? 0.0 : exp(d);
                    }
                    colacc[j] += w; racc += w;
                    const int u2 = h2 / K;
                    if (pathological_u(P, u1, u2)) raccp += w;
                    if (!(dup1 || ((dup2 >> j) & 1))) {
                        raccu += w;
                        if (w >= eps) emit_joint(g, pi, u1, u2, w);
                    }
                    if (ml == M) argkey = min(argkey, ((unsigned long long)h1 << 40) | (unsigned long long)(row + col));
                }
                if (__any_sync(0xffffffffu, racc != 0.0)) {
                    for (int d = 16; d > 0; d >>= 1) {
                        racc += __shfl_down_sync(0xffffffffu, racc, d);
                        raccp += __shfl_down_sync(0xffffffffu, raccp, d);
                        raccu += __shfl_down_sync(0xffffffffu, raccu, d);
                    }
                    if (lane == 0) { ph1[i1] += racc; sum_all += racc; sum_path += raccp; sum_uniq += raccu; }
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) s_col[warp][lane + 32 * j] = colacc[j];
            __syncthreads();
            double cs = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) cs += s_col[w][tid];
            if (cb + tid < n2) R.colpart[(long long)c * n2 + cb + tid] = cs;
            __syncthreads();
        }
        for (int d = 16; d > 0; d >>= 1) argkey = min(argkey, __shfl_xor_sync(0xffffffffu, argkey, d));
        if (lane == 0) { s_sum[warp][0] = sum_all; s_sum[warp][1] = sum_path; s_sum[warp][2] = sum_uniq; s_arg[warp] = argkey; }
        __syncthreads();
        if (tid == 0) {
            double a = 0.0, p = 0.0, u = 0.0;
            unsigned long long k = ~0ULL;
            for (int w = 0; w < 8; ++w) { a += s_sum[w][0]; p += s_sum[w][1]; u += s_sum[w][2]; k = min(k, s_arg[w]); }
            R.partial[4LL * c] = a; R.partial[4LL * c + 1] = p; R.partial[4LL * c + 2] = u;
            if (k != ~0ULL) atomicMin(&B.argkey, k);
            __threadfence();
            s_last = (atomicAdd(&B.done_b, 1u) == (unsigned)(nc - 1));
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            for (int x = tid; x < n2; x += 256) {
                double acc = 0.0;
                for (int r = 0; r < nc; ++r) acc += __ldcg(R.colpart + (long long)r * n2 + x);
                ph2[x] = acc;
            }
            if (tid == 0) {
                double a = 0.0, p = 0.0, u = 0.0;
                for (int r = 0; r < nc; ++r) { a += __ldcg(R.partial + 4LL * r); p += __ldcg(R.partial + 4LL * r + 1); u += __ldcg(R.partial + 4LL * r + 2); }
                const unsigned long long k = *(volatile unsigned long long *)&B.argkey;
                const int npoints = *(volatile int *)&B.npoints;
                tredsw_grid_result r;
                r.max_ml = maxkey ? M : -INFINITY; r.sum_all = a; r.sum_path = p; r.sum_uniq = u;
                const long long idx = (long long)(k & ((1ULL << 40) - 1));
                r.arg_i1 = (npoints && k != ~0ULL) ? (int)(idx / n2) : -1;
                r.arg_i2 = (npoints && k != ~0ULL) ? (int)(idx % n2) : -1;
                r.n_points = npoints; r.pad = 0;
                g.res[pi] = r;
            }
        }
        __syncthreads();
    }
}

// ---- KDE of paired-end lengths (models.py:428-435): see kde.cuh ---------------------------------------
__global__ void __launch_bounds__(KDE_THREADS) pe_kde_kernel(const int32_t *lens, const int64_t *off, int nproblems,
                                                      double *pdf_out) {
    for (int pi = blockIdx.x; pi < nproblems; pi += gridDim.x)
        kde_block(lens + off[pi], (int)(off[pi + 1] - off[pi]), pdf_out + (int64_t)pi * SPAN);
}

}  // namespace

int tredsw_internal_grid(tredsw_ctx *ctx, const tredsw_grid_problem *d_prob, int nproblems,
                         const int32_t *d_ipool, const double *d_dpool, double *d_surface, double *d_marg,
                         tredsw_grid_result *d_res, long long points_hint, int materialise,
                         tredsw_posterior *d_post, long long post_cap, unsigned long long *d_post_cursor,
                         unsigned long long **d_overflow_flag) {
    GridParams g{};
    g.prob = d_prob; g.ipool = d_ipool; g.dpool = d_dpool; g.surface = d_surface; g.marg = d_marg; g.res = d_res;
    g.small_value = exp(-10.0);
    g.log_small = log(g.small_value);
    g.nproblems = nproblems; g.materialise = materialise;
    g.post = d_post; g.post_cap = post_cap; g.post_cursor = d_post_cursor;
    int rc;
    ctx->mark(2);
    // [BigInfo x np | cursor, overflow (256 B) | lists (1 + np ints) | table arena]
    const size_t big_bytes = (((size_t)nproblems * sizeof(BigInfo)) + 255) & ~(size_t)255;
    const size_t list_bytes = (((size_t)nproblems + 1) * sizeof(int) + 255) & ~(size_t)255;
    // table arena, in doubles: 128 MB serve ~250 long-expansion surfaces; a cohort searched with --fullsearch has
    // one large surface per problem, ~(60 + n_target) doubles of tables per candidate allele each.  A surface
    // that finds the arena short of its tables is evaluated point by point (slower, same result); one that
    // cannot even get its reduction scratch is reported through the overflow flag and the call is repeated
    // with a larger arena.
    long long cap = 16LL << 20;
    if (points_hint > SMALL_LIMIT) {
        const long long side = (long long)ceil(sqrt((double)points_hint));
        const long long want = (long long)nproblems * 160 * side;
        if (want > cap) cap = want < (1LL << 30) ? want : (1LL << 30);
    }
    if ((long long)(ctx->d_ftab.cap / sizeof(double)) - (long long)((big_bytes + 256 + list_bytes) / sizeof(double)) > cap)
        cap = (long long)(ctx->d_ftab.cap / sizeof(double)) - (long long)((big_bytes + 256 + list_bytes) / sizeof(double));
    if (ctx->d_ftab.ensure(big_bytes + 256 + list_bytes + (size_t)cap * sizeof(double)) != TREDSW_OK) {
        cudaGetLastError();                                            // not enough memory for the big arena
        cap = 16LL << 20;
        if ((rc = ctx->d_ftab.ensure(big_bytes + 256 + list_bytes + (size_t)cap * sizeof(double)))) return rc;
    }
    unsigned char *base = ctx->d_ftab.as<unsigned char>();
    g.big = reinterpret_cast<BigInfo *>(base);
    g.fcursor = reinterpret_cast<unsigned long long *>(base + big_bytes);
    g.lists = reinterpret_cast<int *>(base + big_bytes + 256);
    g.ftab = reinterpret_cast<double *>(base + big_bytes + 256 + list_bytes);
    g.ftab_cap = cap;
    if (d_overflow_flag) *d_overflow_flag = g.fcursor;                 // [0] doubles needed, [1] overflow
    CUDA_TRY(cudaMemsetAsync(g.fcursor, 0, 256 + sizeof(int), ctx->stream));   // cursor, flag, lists[0]
    grid_small_kernel<<<(nproblems + 7) / 8, 256, 0, ctx->stream>>>(g);
    const int nsetup = ctx->sm_count * 8, nrows = ctx->sm_count * 4;
    grid_big_setup_kernel<<<nsetup, 256, 0, ctx->stream>>>(g);
    grid_rows_eval_kernel<<<nrows, 256, 0, ctx->stream>>>(g);
    grid_rows_reduce_kernel<<<nrows, 256, 0, ctx->stream>>>(g);
    CUDA_TRY(cudaGetLastError());
    ctx->mark(3);
    ctx->launches += 4;
    return TREDSW_OK;
}

extern "C" int tredsw_likelihood_grid(tredsw_ctx *ctx, const tredsw_grid_problem *problems, int32_t nproblems,
                                      const int32_t *ipool, int64_t n_ipool, const double *dpool,
                                      int64_t n_dpool, double *surface, int64_t n_surface, double *marg,
                                      int64_t n_marg, tredsw_grid_result *results, uint32_t flags) {
    if (!ctx) { tredsw_set_error("null context"); return TREDSW_ERR_ARG; }
    if (nproblems < 0 || !problems || !results) { tredsw_set_error("bad arguments"); return TREDSW_ERR_ARG; }
    if (nproblems == 0) return TREDSW_OK;
    const bool dev = dev_ptrs(flags);
    if (dev && !surface) { tredsw_set_error("device mode needs the surface buffer (it is the kernels' scratch)"); return TREDSW_ERR_ARG; }
    std::lock_guard<std::mutex> lock(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    // host mode: the surface is materialised iff the caller wants it back; device mode: unless TREDSW_GRID_NO_SURFACE
    const int materialise = dev ? ((flags & TREDSW_GRID_NO_SURFACE) ? 0 : 1) : ((surface && n_surface > 0) ? 1 : 0);
    GridParams g{};
    int rc;
    long long max_points = 0;
    if (!dev) {
        for (int i = 0; i < nproblems; ++i) {
            const tredsw_grid_problem &P = problems[i];
            if (P.n_h1 < 0 || P.n_h2 < 0 || P.period < 1 || P.readlen < 1 ||
                P.off_surface + (long long)P.n_h1 * P.n_h2 > n_surface || P.off_ph1 + P.n_h1 > n_marg ||
                P.off_ph2 + P.n_h2 > n_marg) { tredsw_set_error("grid problem %d out of range", i); return TREDSW_ERR_ARG; }
            long long t = (long long)P.n_h1 * P.n_h2;
            if (t > max_points) max_points = t;
        }
    } else {
        max_points = n_surface;   // upper bound
    }
    if ((rc = stage_in(ctx, ctx->d_prob, problems, (size_t)nproblems, flags, &g.prob))) return rc;
    if ((rc = stage_in(ctx, ctx->d_ipool, ipool, (size_t)n_ipool, flags, &g.ipool))) return rc;
    if ((rc = stage_in(ctx, ctx->d_dpool, dpool, (size_t)n_dpool, flags, &g.dpool))) return rc;
    if (dev) { g.surface = surface; g.marg = marg; g.res = results; }
    else {
        if ((rc = ctx->d_surface.ensure((size_t)(n_surface > 0 ? n_surface : 1) * sizeof(double)))) return rc;
        if ((rc = ctx->d_marg.ensure((size_t)(n_marg > 0 ? n_marg : 1) * sizeof(double)))) return rc;
        if ((rc = ctx->d_res.ensure((size_t)nproblems * sizeof(tredsw_grid_result)))) return rc;
        g.surface = ctx->d_surface.as<double>(); g.marg = ctx->d_marg.as<double>();
        g.res = ctx->d_res.as<tredsw_grid_result>();
    }
    for (int attempt = 0; attempt < 2; ++attempt) {
        unsigned long long *d_flag = nullptr;
        if ((rc = tredsw_internal_grid(ctx, g.prob, nproblems, g.ipool, g.dpool, g.surface, g.marg, g.res, max_points,
                                       materialise, nullptr, 0, nullptr, &d_flag))) return rc;
        if (dev) return TREDSW_OK;
        unsigned long long h_flag[2] = {0, 0};
        CUDA_TRY(cudaMemcpyAsync(h_flag, d_flag, sizeof(h_flag), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (!h_flag[1]) break;
        if (attempt == 1) { tredsw_set_error("likelihood table arena overflow (%llu doubles needed)", h_flag[0]); return TREDSW_ERR_UNSUPPORTED; }
        const size_t big_bytes = (((size_t)nproblems * sizeof(BigInfo)) + 255) & ~(size_t)255;
        const size_t list_bytes = (((size_t)nproblems + 1) * sizeof(int) + 255) & ~(size_t)255;
        if ((rc = ctx->d_ftab.ensure(big_bytes + 256 + list_bytes + (size_t)(h_flag[0] + h_flag[0] / 8) * sizeof(double)))) return rc;
    }
    if (surface && n_surface > 0)
        CUDA_TRY(cudaMemcpyAsync(surface, g.surface, (size_t)n_surface * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (marg && n_marg > 0)
        CUDA_TRY(cudaMemcpyAsync(marg, g.marg, (size_t)n_marg * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(results, g.res, (size_t)nproblems * sizeof(tredsw_grid_result), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return TREDSW_OK;
}

extern "C" int tredsw_pe_kde(tredsw_ctx *ctx, const int32_t *lens, const int64_t *off, int32_t nproblems,
                             double *pdf_out, uint32_t flags) {
    if (!ctx) { tredsw_set_error("null context"); return TREDSW_ERR_ARG; }
    if (nproblems < 0 || !off || !pdf_out) { tredsw_set_error("bad arguments"); return TREDSW_ERR_ARG; }
    if (nproblems == 0) return TREDSW_OK;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int32_t *d_lens; const int64_t *d_off; double *d_out;
    int rc;
    if (dev_ptrs(flags)) { d_lens = lens; d_off = off; d_out = pdf_out; }
    else {
        if ((rc = stage_in(ctx, ctx->d_ipool, lens, (size_t)off[nproblems], flags, &d_lens))) return rc;
        if ((rc = stage_in(ctx, ctx->d_qoff, off, (size_t)nproblems + 1, flags, &d_off))) return rc;
        if ((rc = ctx->d_dpool.ensure((size_t)nproblems * SPAN * sizeof(double)))) return rc;
        d_out = ctx->d_dpool.as<double>();
    }
    int gb = nproblems > ctx->sm_count * 8 ? ctx->sm_count * 8 : nproblems;
    pe_kde_kernel<<<gb, KDE_THREADS, 0, ctx->stream>>>(d_lens, d_off, nproblems, d_out);
    CUDA_TRY(cudaGetLastError());
    ctx->launches += 1;
    if (!dev_ptrs(flags)) {
        CUDA_TRY(cudaMemcpyAsync(pdf_out, d_out, (size_t)nproblems * SPAN * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    return TREDSW_OK;
}

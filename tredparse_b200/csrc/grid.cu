// grid.cu — the (h1, h2) log-likelihood surface of IntegratedCaller and its reductions, FP64.
//
// Replaces the double loop of tredparse/models.py:260-273 (evaluate_spanning / evaluate_partial /
// evaluate_rept / PEMaxLikModel.evaluate at every candidate pair) and the reductions of
// models.py:277-302, 342-368 (max, arg-max with key (ml, -h1), exp-normalised marginals, PP sums).
//
// The reference materialises length-1000 probability vectors per candidate allele and takes the log of
// whole vectors at every grid point; only the entries at the observed keys are ever used.  Here each
// thread evaluates one grid point from the closed forms of SURVEY.md Appendix B: per observed key one
// mixture + log, a Poisson log-pmf, and per spanning pair one gather from the shifted KDE.
//   PS(h)[k]  spanning pdf (models.py:149-168, quirks Q6/Q7)     PT(h)[k] partial pdf (:170-180, Q8)
//   alpha     mixing weights (:182-190)                           R(h)[x]  rolled PE pdf (:441-458)
// Arithmetic order follows the reference (sum over keys in the order given, ml1+ml2+ml3+ml4).
#include "internal.cuh"
#include "kde.cuh"
#include <math.h>
#include <cooperative_groups.h>

namespace {

constexpr int SPAN = 1000;
constexpr int NSTEP = 37;
constexpr int DEV = 18;

struct GridParams {
    const tredsw_grid_problem *prob;
    const int32_t *ipool;
    const double *dpool;
    double *surface;
    double *marg;
    tredsw_grid_result *res;
    double small_value, really_small, log_small;
    // per-problem constants and far-region tables of the medium / large surfaces (grid_setup / grid_fill)
    struct FarInfo *far;            // [nproblems]
    double *ftab;                   // arena of per-problem tables
    long long ftab_cap;
    unsigned long long *fcursor;    // arena cursor
};

// Where the longer allele h2 lies beyond every observed key, the partial clamp and the read length
// (list index i2 >= fam, the MID region), the spanning and partial terms depend on h1 only; where it is also
// shifted past the KDE support (i2 >= fa2, the FAR region) so does the paired-end term; and the repeat-only
// term is a function of dsum = max(h1-L,1) + max(h2-L,1) everywhere.  With the per-row table
// rows[i1] = {span(h1) + partial(h1), pe(h1)} and rept[dsum - 2], a far point is
//   ml = ((rows[i1].c12) + rept[dsum-2]) + rows[i1].pe
// and a mid point the same with the paired-end term evaluated directly — the same additions in the same order
// as the general evaluation, on operands produced by the same device functions, hence the same bits: two
// table reads instead of tens of logarithms.  On a --fullsearch / long-expansion grid (10^5 - 10^6 points)
// > 90 % of the points are far, most of the rest mid.
// Where the paired-end term is genuinely two-dimensional (i2 < fa2) its operands come from two more tables,
// R1[t][i1] = rolled pdf of allele h1s[i1] at target length t and R2[t][i2] likewise for h2s[i2] (the values
// pe_roll returns), so that a point costs two coalesced loads, the mixture and the logarithm per target pair.
struct FarInfo {
    unsigned long long maxkey;   // ordered key of the surface maximum (atomicMax of the tiles kernel); 0 = no point
    long long off;               // ftab: rows[2 * n_h1] ({c12, pe} pairs), then rept[nd]
    double lgamma_k1;            // lgamma(n_rept + 1) of the Poisson term
    double sig_mp;               // sigma(max_partial)
    int tmin;                    // smallest pair length >= MINPE (after numpy's negative-index wrap); INT_MAX if none
    int fam, fa2;                // first h2 index of the mid / far region
    int ok;                      // tables present
    int sorted;                  // both candidate lists are non-decreasing (lets the reduction skip h1 > h2 chunks)
    int nd;                      // entries of rept[]
    int hrep;                    // a far allele (the largest h2)
    int npe;                     // paired-end tables present: R1[npe][n_h1], R2[npe][fa2] after rept[] (npe = n_target)
};

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
    return idx == DEV ? (1.0 - sig) : step[idx] * sig;
}

__device__ __forceinline__ double pdf_part(const double *step, int hc, double sig_c, double c, int k) {
    double v = (k < hc) ? c : 0.0;
    return v + c * pdf_span(step, hc, sig_c, k);
}

__device__ __forceinline__ double pe_roll(const double *pdf, int h, int ref, int minpe, int x, double eps) {
    if (x < minpe) return eps;
    const int y = x + h - ref;
    if (y < 0 || y >= SPAN) return eps;
    return pdf[y];
}

// log(max(v, eps)) with a one-entry memo: consecutive keys very often give bit-identical mixtures (every
// partial key far below both alleles sees the same alpha*c1 + (1-alpha)*c2; every pair length shifted off
// the KDE support sees eps), and log of the same double is the same double — so the surface is unchanged
// while most of the FP64 log evaluations of a large (h1, h2) grid disappear.
struct LogMemo {
    double v, l;
    __device__ __forceinline__ double operator()(double x, double eps, double log_small) {
        if (x != v) { v = x; l = (x < eps) ? log_small : log(x); }
        return l;
    }
};

// Per-problem values shared by all points of a tile (computed once per tile by one thread).
struct TileShared {
    double lgamma_k1;      // lgamma(n_rept + 1) of the Poisson term
    double sig_mp;         // sigma(max_partial): the stutter probability of every allele clamped to max_partial
    int tmin;              // smallest pair length >= MINPE (after numpy's negative-index wrap); INT_MAX if none
};

// One grid point.  The arithmetic (operation order, rounding) is exactly that of the straightforward loops
// over all keys; the shortcuts only skip work whose result is known in advance:
//   * pdf_span(h)[k] = 0 unless h-18 <= k <= h+18, so a partial key below both clamped alleles by more than
//     18 sees the mixture alpha*c1 + (1-alpha)*c2 and one above both sees 0 — no pdf evaluation, no sigma;
//   * a spanning key farther than 18 from both alleles sees 0;
//   * when both alleles shift every pair length off the KDE support, every pair sees eps.
// On the large grids of --fullsearch / long-expansion searches almost all points are of that kind.
__device__ __forceinline__ double ml_span_term(const tredsw_grid_problem &P, const GridParams &g, int h1, int h2) {
    if (P.n_span <= 0) return 0.0;
    const int32_t *skey = g.ipool + P.off_span, *scnt = skey + P.n_span;
    const double *step = g.dpool + P.off_step;
    const double eps = g.small_value;
    const int t2 = P.readlen - 18;
    const int s1 = max(0, t2 - h1), s2 = max(0, t2 - h2);
    const double a = (s1 + s2) ? (double)s1 * 1.0 / (double)(s1 + s2) : 0.5;
    const int lo = min(h1, h2) - DEV, hi = max(h1, h2) + DEV;
    double sg1 = 0.0, sg2 = 0.0;
    bool have_sigma = false;
    double acc = 0.0;
    LogMemo lg{-1.0, 0.0};
    for (int i = 0; i < P.n_span; ++i) {
        const int k = skey[i];
        double v;
        if (k < lo || k > hi) v = 0.0;          // == a * 0 + (1 - a) * 0
        else {
            if (!have_sigma) { sg1 = sigma_h(P, h1); sg2 = sigma_h(P, h2); have_sigma = true; }
            const double p1 = pdf_span(step, h1, sg1, k), p2 = pdf_span(step, h2, sg2, k);
            v = __dadd_rn(__dmul_rn(a, p1), __dmul_rn(1.0 - a, p2));
        }
        double l = lg(v, eps, g.log_small);
        acc = __dadd_rn(acc, __dmul_rn(l, (double)scnt[i]));
    }
    return acc;
}

__device__ __forceinline__ double ml_part_term(const tredsw_grid_problem &P, const GridParams &g, int h1, int h2, double sig_mp) {
    if (P.n_part <= 0) return 0.0;
    const int32_t *pkey = g.ipool + P.off_part, *pcnt = pkey + P.n_part;
    const double *step = g.dpool + P.off_step;
    const double eps = g.small_value;
    const int t1 = P.readlen - 9;
    const int s1 = min(h1, t1), s2 = min(h2, t1);
    const double a = (s1 + s2) ? (double)s1 * 1.0 / (double)(s1 + s2) : 0.5;
    const int hc1 = min(h1, P.max_partial), hc2 = min(h2, P.max_partial);
    const double c1 = 1.0 / (double)(hc1 + 1), c2 = 1.0 / (double)(hc2 + 1);
    const int lo = min(hc1, hc2) - DEV, hi = max(hc1, hc2) + DEV;
    const double v_bulk = __dadd_rn(__dmul_rn(a, c1), __dmul_rn(1.0 - a, c2));   // p1 = c1 + c1 * 0, p2 = c2 + c2 * 0
    double sg1 = 0.0, sg2 = 0.0;
    bool have_sigma = false;
    double acc = 0.0;
    LogMemo lg{-1.0, 0.0};
    for (int i = 0; i < P.n_part; ++i) {
        const int k = pkey[i];
        double v;
        if (k < lo) v = v_bulk;
        else if (k > hi) v = 0.0;
        else {
            if (!have_sigma) {
                sg1 = hc1 == P.max_partial ? sig_mp : sigma_h(P, hc1);
                sg2 = hc2 == P.max_partial ? sig_mp : sigma_h(P, hc2);
                have_sigma = true;
            }
            const double p1 = pdf_part(step, hc1, sg1, c1, k), p2 = pdf_part(step, hc2, sg2, c2, k);
            v = __dadd_rn(__dmul_rn(a, p1), __dmul_rn(1.0 - a, p2));
        }
        double l = lg(v, eps, g.log_small);
        acc = __dadd_rn(acc, __dmul_rn(l, (double)pcnt[i]));
    }
    return acc;
}

// repeat-only reads: Poisson (scipy: exp(xlogy(k, mu) - gammaln(k + 1) - mu)); dsum = max(h1-L,1) + max(h2-L,1)
__device__ __forceinline__ double ml_rept_term(const tredsw_grid_problem &P, int dsum, double lgamma_k1) {
    const double mu = (double)dsum * P.half_depth / (double)P.readlen;
    const double kk = (double)P.n_rept;
    const double xl = (P.n_rept == 0) ? 0.0 : __dmul_rn(kk, log(mu));
    const double pk = __dsub_rn(__dsub_rn(xl, lgamma_k1), mu);    // no FMA contraction: table and direct path agree
    // log(max(exp(pk), e^-100)): log(exp(pk)) is pk to within an ulp of the pmf (~1e-16 absolute on a term
    // of magnitude 0.1..100, far inside the 1e-9 relative bar) — two transcendentals less per point
    return pk > -100.0 ? pk : -100.0;
}

__device__ __forceinline__ double ml_pe_term(const tredsw_grid_problem &P, const GridParams &g, int h1, int h2, int tmin) {
    if (!P.run_pe) return 0.0;
    const double eps = g.small_value;
    const double *pdf = g.dpool + P.off_pdf;
    const int32_t *tl = g.ipool + P.off_target;
    double acc = 0.0;
    const long long off1 = (long long)h1 - P.pe_ref, off2 = (long long)h2 - P.pe_ref;
    if (tmin == 0x7fffffff || (tmin + off1 >= SPAN && tmin + off2 >= SPAN)) {
        // every pair length is below MINPE or shifted past the end of the support: 0.5*eps + 0.5*eps = eps
        const double l = log(eps);
        for (int i = 0; i < P.n_target; ++i) acc = __dadd_rn(acc, l);
    } else {
        LogMemo lg{-1.0, 0.0};
        for (int i = 0; i < P.n_target; ++i) {
            int x = tl[i];
            if (x < 0) x += SPAN;                   // numpy negative-index wrap (models.py:473)
            const double r1 = pe_roll(pdf, h1, P.pe_ref, P.pe_minpe, x, eps);
            const double r2 = pe_roll(pdf, h2, P.pe_ref, P.pe_minpe, x, eps);
            double v = __dadd_rn(__dmul_rn(0.5, r1), __dmul_rn(0.5, r2));
            double l = lg(v, eps, g.log_small);
            acc = __dadd_rn(acc, l);
        }
    }
    return acc;
}

// ml_pe_term from the tabulated operands (FarInfo.npe > 0, i2 < fa2): same mixture, same memo, same sum
__device__ __forceinline__ double ml_pe_tab(const double *R1, const double *R2, int n1, int n2, int npe, int i1, int i2,
                                            double eps, double log_small) {
    double acc = 0.0;
    LogMemo lg{-1.0, 0.0};
    R1 += i1; R2 += i2;
    for (int t = 0; t < npe; ++t) {
        const double v = __dadd_rn(__dmul_rn(0.5, R1[(long long)t * n1]), __dmul_rn(0.5, R2[(long long)t * n2]));
        acc = __dadd_rn(acc, lg(v, eps, log_small));
    }
    return acc;
}

__device__ double point_ml(const tredsw_grid_problem &P, const GridParams &g, int h1, int h2, const TileShared &T) {
    double ml = ml_span_term(P, g, h1, h2);
    ml = __dadd_rn(ml, ml_part_term(P, g, h1, h2, T.sig_mp));
    ml = __dadd_rn(ml, ml_rept_term(P, max(h1 - P.readlen, 1) + max(h2 - P.readlen, 1), T.lgamma_k1));
    ml = __dadd_rn(ml, ml_pe_term(P, g, h1, h2, T.tmin));
    return ml;
}

// The surfaces of a batch differ in size by orders of magnitude (a handful of points for a locus with
// spanning reads only, 10^4-10^6 when the h2 range is extended or with --fullsearch), so the points of all
// problems are flattened into tiles of GRID_TILE points: tile_start = exclusive prefix sum of the tiles per
// problem (device scan), then a persistent kernel strides over the tiles.
constexpr int GRID_TILE = 256;
constexpr int TILE_PTS = 4;                          // points per thread of grid_surface_tiles_kernel
constexpr int TILE_POINTS = GRID_TILE * TILE_PTS;    // points per tile of a medium / large surface
constexpr int GRID_CLUSTER = 8;       // CTAs per problem in the cluster variant of the reduction

// Problem classes of the reductions, by number of surface points.
constexpr long long GRID_WARP_LIMIT = 2048;        // <= : one warp per problem
constexpr long long GRID_BLOCK_LIMIT = 16384;      // <= : one CTA per problem;  > : a cluster of 8 CTAs

// Single block: pt_start = exclusive prefix sum of the points per problem (pt_start[np] = all points), and the
// index lists of the "block" and "cluster" class problems in ascending order (lists[0] = #block, lists[1] =
// #cluster, then the block list at lists + 2, the cluster list at lists + 2 + np).
__global__ void __launch_bounds__(1024) grid_tiles_kernel(const tredsw_grid_problem *prob, int nproblems,
                                                          long long *pt_start, long long *tile_start, int *lists) {
    __shared__ long long part[1024], tpart[1024];
    __shared__ int pb[1024], pc[1024];
    const int tid = threadIdx.x;
    const int chunk = (nproblems + 1023) / 1024;
    const int lo = min(nproblems, tid * chunk), hi = min(nproblems, lo + chunk);
    auto points_of = [&](int i) { return prob[i].n_h2 > 0 ? (long long)prob[i].n_h1 * prob[i].n_h2 : 0LL; };
    // small surfaces (<= GRID_WARP_LIMIT points) are concatenated point by point (pt_start), the others are
    // cut into their own tiles of GRID_TILE points (tile_start)
    long long s = 0, ts = 0;
    int nb = 0, nc = 0;
    for (int i = lo; i < hi; ++i) {
        const long long t = points_of(i);
        if (t > GRID_WARP_LIMIT) ts += (t + TILE_POINTS - 1) / TILE_POINTS; else s += t;
        if (t > GRID_BLOCK_LIMIT) ++nc; else if (t > GRID_WARP_LIMIT) ++nb;
    }
    part[tid] = s; tpart[tid] = ts; pb[tid] = nb; pc[tid] = nc;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {               // inclusive Hillis-Steele scans of the partial sums
        const long long v = tid >= d ? part[tid - d] : 0, tv = tid >= d ? tpart[tid - d] : 0;
        const int vb = tid >= d ? pb[tid - d] : 0, vc = tid >= d ? pc[tid - d] : 0;
        __syncthreads();
        part[tid] += v; tpart[tid] += tv; pb[tid] += vb; pc[tid] += vc;
        __syncthreads();
    }
    long long run = part[tid] - s, trun = tpart[tid] - ts;
    int wb = pb[tid] - nb, wc = pc[tid] - nc;
    for (int i = lo; i < hi; ++i) {
        const long long t = points_of(i);
        pt_start[i] = run; tile_start[i] = trun;
        if (t > GRID_WARP_LIMIT) trun += (t + TILE_POINTS - 1) / TILE_POINTS; else run += t;
        if (t > GRID_BLOCK_LIMIT) lists[2 + nproblems + wc++] = i; else if (t > GRID_WARP_LIMIT) lists[2 + wb++] = i;
    }
    if (tid == 1023) { pt_start[nproblems] = part[1023]; tile_start[nproblems] = tpart[1023]; lists[0] = pb[1023]; lists[1] = pc[1023]; }
}

__device__ __forceinline__ TileShared tile_shared_of(const GridParams &g, const tredsw_grid_problem &Q) {
    TileShared T;
    T.lgamma_k1 = lgamma((double)Q.n_rept + 1.0);
    T.sig_mp = sigma_h(Q, Q.max_partial);
    int tmin = 0x7fffffff;
    if (Q.run_pe) {
        const int32_t *tl = g.ipool + Q.off_target;
        for (int i = 0; i < Q.n_target; ++i) { int x = tl[i]; if (x < 0) x += SPAN; if (x >= Q.pe_minpe && x < tmin) tmin = x; }
    }
    T.tmin = tmin;
    return T;
}

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

// One block per medium / large surface (class lists of grid_tiles_kernel): per-problem constants, the
// mid / far thresholds and the table allocation (FarInfo).  All scans are block-parallel.
constexpr long long FAR_MIN_POINTS = 4096;     // tables only pay for at least this many mid + far points
constexpr unsigned long long PE_TAB_MAX = 1ULL << 20;   // doubles per problem for the paired-end tables
__global__ void __launch_bounds__(256) grid_setup_kernel(GridParams g, const int *lists, int nproblems) {
    __shared__ int s8[8];
    const int nb = lists[0], nc = lists[1];
    const int tid = threadIdx.x;
    for (int k = blockIdx.x; k < nb + nc; k += gridDim.x) {
        const int pi = k < nb ? lists[2 + k] : lists[2 + nproblems + (k - nb)];
        const tredsw_grid_problem &P = g.prob[pi];
        const int32_t *h1s = g.ipool + P.off_h1, *h2s = g.ipool + P.off_h2;
        const int32_t *skey = g.ipool + P.off_span, *tl = g.ipool + P.off_target;
        const bool two = P.ploidy != 1 && P.n_h2 > 0;
        int ks = -1000000, ntmin = -0x7fffffff, mx1 = -0x7fffffff, mx2 = -0x7fffffff, unsorted = 0;
        for (int i = tid; i < P.n_span; i += 256) ks = max(ks, skey[i]);
        if (P.run_pe)
            for (int i = tid; i < P.n_target; i += 256) { int x = tl[i]; if (x < 0) x += SPAN; if (x >= P.pe_minpe) ntmin = max(ntmin, -x); }
        for (int i = tid; i < P.n_h1; i += 256) { mx1 = max(mx1, h1s[i]); if (i > 0 && h1s[i - 1] > h1s[i]) unsorted = 1; }
        if (two)
            for (int i = tid; i < P.n_h2; i += 256) { mx2 = max(mx2, h2s[i]); if (i > 0 && h2s[i - 1] > h2s[i]) unsorted = 1; }
        ks = block_max_int(ks, s8); ntmin = block_max_int(ntmin, s8);
        mx1 = block_max_int(mx1, s8); mx2 = block_max_int(mx2, s8); unsorted = block_max_int(unsorted, s8);
        const int tmin = ntmin == -0x7fffffff ? 0x7fffffff : -ntmin;
        const int t1 = P.readlen - 9;
        // h2 >= H1: no spanning key within 18, partial clamp and mixing weights saturated, h2 > readlen
        const int H1 = max(max(ks + DEV + 1, P.max_partial), max(t1, P.readlen + 1));
        // h2 >= H2: every pair length >= MINPE is shifted past the support as well
        long long H2 = H1;
        if (P.run_pe && tmin != 0x7fffffff) H2 = max((long long)H1, (long long)P.pe_ref + SPAN - tmin);
        int l1 = -1, l2 = -1;                          // last h2 index below H1 / H2
        if (two)
            for (int i = tid; i < P.n_h2; i += 256) { const int h = h2s[i]; if (h < H1) l1 = max(l1, i); if (h < H2) l2 = max(l2, i); }
        l1 = block_max_int(l1, s8); l2 = block_max_int(l2, s8);
        if (tid == 0) {
            FarInfo f;
            f.maxkey = 0; f.off = 0; f.ok = 0; f.npe = 0;
            f.lgamma_k1 = lgamma((double)P.n_rept + 1.0);
            f.sig_mp = sigma_h(P, P.max_partial);
            f.tmin = tmin;
            f.fam = two ? l1 + 1 : P.n_h2; f.fa2 = two ? l2 + 1 : P.n_h2;
            f.sorted = unsorted ? 0 : 1;
            f.hrep = mx2;
            f.nd = (two && P.n_h1 > 0) ? max(mx1 - P.readlen, 1) + max(mx2 - P.readlen, 1) - 1 : 0;
            if (two && P.n_h1 > 0 && (long long)(P.n_h2 - f.fam) * P.n_h1 >= FAR_MIN_POINTS) {
                const unsigned long long base = (unsigned long long)(2LL * P.n_h1 + f.nd);
                const unsigned long long pe = (P.run_pe && P.n_target > 0) ? (unsigned long long)P.n_target * (unsigned long long)(P.n_h1 + f.fa2) : 0ULL;
                const bool want_pe = pe > 0 && pe <= PE_TAB_MAX;
                const unsigned long long need = (base + (want_pe ? pe : 0ULL) + 1ULL) & ~1ULL;   // even: rows stay 16-byte aligned
                const unsigned long long off = atomicAdd(g.fcursor, need);
                if ((long long)(off + need) <= g.ftab_cap) { f.off = (long long)off; f.ok = 1; f.npe = want_pe ? P.n_target : 0; }
            }
            g.far[pi] = f;
        }
    }
}

// Fill the tables: FILL_SPLIT blocks per problem over the rows and the dsum entries.
constexpr int FILL_SPLIT = 8;
__global__ void __launch_bounds__(256) grid_fill_kernel(GridParams g, const int *lists, int nproblems) {
    const int nb = lists[0], nc = lists[1];
    for (int kk = blockIdx.x; kk < (nb + nc) * FILL_SPLIT; kk += gridDim.x) {
        const int k = kk / FILL_SPLIT, part = kk - k * FILL_SPLIT;
        const int pi = k < nb ? lists[2 + k] : lists[2 + nproblems + (k - nb)];
        const FarInfo &F = g.far[pi];
        if (!F.ok) continue;
        const tredsw_grid_problem &P = g.prob[pi];
        const int32_t *h1s = g.ipool + P.off_h1;
        double *rows = g.ftab + F.off, *rept = rows + 2 * (long long)P.n_h1;
        const int hrep = F.hrep, tmin = F.tmin;
        const double sig_mp = F.sig_mp, lgk = F.lgamma_k1;
        const int nbase = P.n_h1 + F.nd, n1 = F.npe * P.n_h1, n = nbase + n1 + F.npe * F.fa2;
        const int32_t *h2s = g.ipool + P.off_h2, *tl = g.ipool + P.off_target;
        const double *pdf = g.dpool + P.off_pdf;
        double *R1 = rept + F.nd, *R2 = R1 + n1;
        for (int e = part * 256 + threadIdx.x; e < n; e += 256 * FILL_SPLIT) {
            if (e < P.n_h1) {
                const int h1 = h1s[e];
                rows[2 * e] = __dadd_rn(ml_span_term(P, g, h1, hrep), ml_part_term(P, g, h1, hrep, sig_mp));
                rows[2 * e + 1] = ml_pe_term(P, g, h1, hrep, tmin);
            } else if (e < nbase) {
                rept[e - P.n_h1] = ml_rept_term(P, e - P.n_h1 + 2, lgk);
            } else {
                // operands of the two-dimensional paired-end term, exactly as ml_pe_term obtains them
                int idx = e - nbase;
                const bool first = idx < n1;
                if (!first) idx -= n1;
                const int width = first ? P.n_h1 : F.fa2;
                const int t = idx / width, i = idx - t * width;
                int x = tl[t];
                if (x < 0) x += SPAN;                   // numpy negative-index wrap (models.py:473)
                (first ? R1 : R2)[idx] = pe_roll(pdf, first ? h1s[i] : h2s[i], P.pe_ref, P.pe_minpe, x, g.small_value);
            }
        }
    }
}

// Large and medium surfaces: persistent over their tiles of TILE_POINTS consecutive points (4 per thread, 256
// apart: the per-tile work — tile -> problem search, index division, max reduction — is shared by 1024 points;
// most points of a large surface are table look-ups, so that overhead is what they cost).  Besides the
// surface, every tile contributes to the maximum of its problem (ordered-key atomicMax — order independent,
// hence deterministic), so that the reduction needs a single pass.
__global__ void __launch_bounds__(GRID_TILE, 4) grid_surface_tiles_kernel(GridParams g, int nproblems, const long long *tile_start) {
    const long long ntiles = tile_start[nproblems];
    __shared__ int s_pi;
    __shared__ unsigned long long s_wmax[GRID_TILE / 32];
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        if (threadIdx.x == 0) {                        // tile -> problem: last p with tile_start[p] <= tile
            int lo = 0, hi = nproblems;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (tile_start[mid] <= tile) lo = mid; else hi = mid;
            }
            s_pi = lo;
        }
        __syncthreads();
        const int pi = s_pi;
        const tredsw_grid_problem &P = g.prob[pi];
        const FarInfo &F = g.far[pi];
        const int n_h1 = P.n_h1, n_h2 = P.n_h2, readlen = P.readlen;
        const bool haploid = P.ploidy == 1;
        const long long total = (long long)n_h1 * n_h2;
        const int32_t *h1s = g.ipool + P.off_h1, *h2s = g.ipool + P.off_h2;
        double *surf = g.surface + P.off_surface;
        const int f_ok = F.ok, fam = F.fam, fa2 = F.fa2, npe = F.npe;
        const double *rows = g.ftab + F.off, *rept = rows + 2 * (long long)n_h1, *R1 = rept + F.nd;
        TileShared T;
        T.lgamma_k1 = F.lgamma_k1; T.sig_mp = F.sig_mp; T.tmin = F.tmin;
        long long t = (tile - tile_start[pi]) * TILE_POINTS + threadIdx.x;
        int i1 = 0, i2 = 0;
        if (t < total) {
            if (total < 0x7fffffffLL) { i1 = (int)((unsigned)t / (unsigned)n_h2); i2 = (int)((unsigned)t - (unsigned)i1 * (unsigned)n_h2); }
            else { i1 = (int)(t / n_h2); i2 = (int)(t % n_h2); }
        }
        unsigned long long key = 0;
#pragma unroll 1
        for (int k = 0; k < TILE_PTS && t < total; ++k, t += GRID_TILE) {
            const int h1 = h1s[i1];
            const int h2 = haploid ? h1 : h2s[i2];
            double ml = -INFINITY;
            if (h1 <= h2) {
                if (f_ok) {
                    const double rp = rept[max(h1 - readlen, 1) + max(h2 - readlen, 1) - 2];
                    const double2 row = reinterpret_cast<const double2 *>(rows)[i1];
                    double pe;
                    if (i2 >= fa2) pe = row.y;
                    else if (npe > 0) pe = ml_pe_tab(R1, R1 + (long long)npe * n_h1, n_h1, fa2, npe, i1, i2, g.small_value, g.log_small);
                    else pe = ml_pe_term(P, g, h1, h2, T.tmin);
                    if (i2 >= fam) {
                        ml = __dadd_rn(__dadd_rn(row.x, rp), pe);
                    } else {
                        ml = __dadd_rn(ml_span_term(P, g, h1, h2), ml_part_term(P, g, h1, h2, T.sig_mp));
                        ml = __dadd_rn(ml, rp);
                        ml = __dadd_rn(ml, pe);
                    }
                } else {
                    ml = point_ml(P, g, h1, h2, T);
                }
                key = max(key, ord_key(ml));
            }
            surf[t] = ml;
            // next point of this thread: GRID_TILE further along the row-major order
            if (n_h2 >= GRID_TILE) { i2 += GRID_TILE; if (i2 >= n_h2) { i2 -= n_h2; ++i1; } }
            else { const int adv = i2 + GRID_TILE; const int q = adv / n_h2; i1 += q; i2 = adv - q * n_h2; }
        }
        key = max(key, __shfl_xor_sync(0xffffffffu, key, 16));
        key = max(key, __shfl_xor_sync(0xffffffffu, key, 8));
        key = max(key, __shfl_xor_sync(0xffffffffu, key, 4));
        key = max(key, __shfl_xor_sync(0xffffffffu, key, 2));
        key = max(key, __shfl_xor_sync(0xffffffffu, key, 1));
        if ((threadIdx.x & 31) == 0) s_wmax[threadIdx.x >> 5] = key;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long m = s_wmax[0];
#pragma unroll
            for (int w = 1; w < GRID_TILE / 32; ++w) m = max(m, s_wmax[w]);
            unsigned long long *dst = &g.far[pi].maxkey;
            if (m != 0 && m > *(volatile unsigned long long *)dst) atomicMax(dst, m);
        }
    }
}

// Small surfaces: persistent over tiles of GRID_TILE consecutive points of their CONCATENATION, every thread
// finds its own problem and computes the per-problem values itself — thousands of 10-point surfaces cost a few
// tiles, not one tile each.
__global__ void __launch_bounds__(GRID_TILE) grid_surface_points_kernel(GridParams g, int nproblems, const long long *pt_start) {
    const long long npoints = pt_start[nproblems];
    for (long long gidx = (long long)blockIdx.x * GRID_TILE + threadIdx.x; gidx < npoints; gidx += (long long)gridDim.x * GRID_TILE) {
        int lo = 0, hi = nproblems;        // last p with pt_start[p] <= gidx (problems without points here share
        while (hi - lo > 1) {              // their start with the next one; the last of equals is the owner)
            const int mid = (lo + hi) >> 1;
            if (pt_start[mid] <= gidx) lo = mid; else hi = mid;
        }
        const tredsw_grid_problem &P = g.prob[lo];
        const TileShared T = tile_shared_of(g, P);
        const int t = (int)(gidx - pt_start[lo]);
        const int32_t *h1s = g.ipool + P.off_h1, *h2s = g.ipool + P.off_h2;
        const int i1 = t / P.n_h2, i2 = t - i1 * P.n_h2;
        const int h1 = h1s[i1];
        const int h2 = (P.ploidy == 1) ? h1 : h2s[i2];
        double ml = -INFINITY;
        if (h1 <= h2) ml = point_ml(P, g, h1, h2, T);
        g.surface[P.off_surface + t] = ml;
    }
}

struct ArgMax { double ml; int h1; long long idx; };
__device__ __forceinline__ bool better(const ArgMax &a, const ArgMax &b) {   // a beats b (Q10)
    if (a.ml != b.ml) return a.ml > b.ml;
    if (a.h1 != b.h1) return a.h1 < b.h1;
    return a.idx < b.idx;
}

// (u1, u2 = h1 / period, h2 / period: division is monotone, so min / max commute with it)
__device__ __forceinline__ bool pathological_u(const tredsw_grid_problem &P, int u1, int u2) {
    const int lo = min(u1, u2), hi = max(u1, u2);
    if (P.expansion) return P.recessive ? (lo >= P.cutoff_risk) : (hi >= P.cutoff_risk);
    return P.recessive ? (hi <= P.cutoff_risk) : (lo <= P.cutoff_risk);
}
__device__ __forceinline__ bool pathological(const tredsw_grid_problem &P, int h1, int h2) {
    const int lo = min(h1, h2) / P.period, hi = max(h1, h2) / P.period;
    if (P.expansion) return P.recessive ? (lo >= P.cutoff_risk) : (hi >= P.cutoff_risk);
    return P.recessive ? (hi <= P.cutoff_risk) : (lo <= P.cutoff_risk);
}

// Reductions of one problem's surface (models.py:277-302, 342-368): arg-max with key (ml, -h1, order) (Q10),
// number of evaluated points, the exp-normalised marginals P_h1 / P_h2 and the PP sums — in ONE pass over the
// surface (its maximum is already known from the tiles kernel).  CS = 1: one CTA per problem (medium
// surfaces).  CS = 8: a thread-block CLUSTER of 8 CTAs per problem for the large surfaces (extended ranges,
// --fullsearch: 10^4 - 10^6 points).  Rows are dealt to the warps of the CTA / cluster; a warp walks its rows
// in chunks of 256 columns, 8 columns per lane, keeps the column partial sums in registers and reduces the row
// sums by shuffles; the column partials of the warps meet in shared memory, those of the CTAs of a cluster in
// distributed shared memory, always in rank order — deterministic.  Chunks of a row that lie entirely in the
// h1 > h2 half are skipped when the candidate lists are sorted (FarInfo.sorted), and exp(ml - max) is not
// evaluated where it is exactly 0 (ml - max < -746), which is almost everywhere on a large surface.
constexpr int RED_CHUNK = 256;        // columns per chunk (8 per lane)
constexpr int RED_SUPER = 2048;       // columns per exchange through (distributed) shared memory
template <int CS>
__global__ void __launch_bounds__(256, 3) grid_reduce_kernel(GridParams g, const int *list, const int *nlist) {
    namespace cg = cooperative_groups;
    __shared__ double s_col[8][RED_CHUNK];
    __shared__ double c_col[CS > 1 ? RED_SUPER : 1];   // this CTA's column partials, read by the cluster
    __shared__ ArgMax s_best[8];
    __shared__ double s_sum[8], s_path[8];
    __shared__ int s_cnt[8];
    __shared__ ArgMax c_best;          // this CTA's partial results, read by the other CTAs of the cluster
    __shared__ int c_cnt;
    __shared__ double c_sum, c_path;
    unsigned crank = 0;
    if (CS > 1) crank = cg::this_cluster().block_rank();
    const int n = *nlist;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // persistent over the problems of this class (the CTAs of a cluster walk the list together)
    for (int k = blockIdx.x / CS; k < n; k += (int)gridDim.x / CS) {
    const int pi = list[k];
    const tredsw_grid_problem &P = g.prob[pi];
    const int32_t *h1s = g.ipool + P.off_h1, *h2s = g.ipool + P.off_h2;
    const double *surf = g.surface + P.off_surface;
    double *ph1 = g.marg + P.off_ph1, *ph2 = g.marg + P.off_ph2;
    const unsigned long long maxkey = g.far[pi].maxkey;
    const double top_ml = maxkey ? ord_val(maxkey) : -INFINITY;
    const bool sorted = g.far[pi].sorted != 0 && P.ploidy != 1;
    const int row0 = (int)crank * 8 + warp, rstep = 8 * CS;
    for (int i1 = row0; i1 < P.n_h1; i1 += rstep) if (lane == 0) ph1[i1] = 0.0;
    ArgMax best{-INFINITY, 0x7fffffff, 0x7fffffffffffffffLL};
    int cnt = 0;
    double sum_all = 0.0, sum_path = 0.0;              // lane 0 of every warp
    for (int sb = 0; sb < P.n_h2; sb += RED_SUPER) {
        const int sb_end = min(sb + RED_SUPER, P.n_h2);
        for (int cb = sb; cb < sb_end; cb += RED_CHUNK) {
            const int h2_last = sorted ? h2s[min(cb + RED_CHUNK, P.n_h2) - 1] : 0x7fffffff;
            int h2c[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { const int c = cb + lane + 32 * j; h2c[j] = (P.ploidy != 1 && c < P.n_h2) ? h2s[c] / P.period : 0; }   // units
            double col[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) col[j] = 0.0;
            for (int i1 = row0; i1 < P.n_h1; i1 += rstep) {
                const int h1 = h1s[i1];
                if (h2_last < h1) continue;                               // the whole chunk has h1 > h2: not evaluated
                const int u1 = h1 / P.period;
                const long long row = (long long)i1 * P.n_h2;
                double v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) { const int c = cb + lane + 32 * j; v[j] = c < P.n_h2 ? surf[row + c] : -INFINITY; }
                double racc = 0.0, raccp = 0.0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const double ml = v[j];
                    if (ml == -INFINITY) continue;                        // not evaluated (h1 > h2) or past the row
                    ++cnt;
                    const double d = ml - top_ml;
                    if (d < -746.0) continue;                             // exp(d) == 0 exactly
                    const double w = exp(d);
                    col[j] += w; racc += w;
                    if (pathological_u(P, u1, (P.ploidy == 1) ? u1 : h2c[j])) raccp += w;
                    if (d == 0.0) {
                        ArgMax c{ml, h1, row + cb + lane + 32 * j};
                        if (better(c, best)) best = c;
                    }
                }
                if (__any_sync(0xffffffffu, racc != 0.0)) {
                    for (int dd = 16; dd > 0; dd >>= 1) {
                        racc += __shfl_down_sync(0xffffffffu, racc, dd);
                        raccp += __shfl_down_sync(0xffffffffu, raccp, dd);
                    }
                    if (lane == 0) { ph1[i1] += racc; sum_all += racc; sum_path += raccp; }
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) s_col[warp][lane + 32 * j] = col[j];
            __syncthreads();
            double c = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) c += s_col[w][tid];
            if (CS > 1) c_col[cb - sb + tid] = c;
            else if (cb + tid < P.n_h2) ph2[cb + tid] = c;
            __syncthreads();
        }
        if (CS > 1) {
            cg::cluster_group cluster = cg::this_cluster();
            cluster.sync();
            for (int x = (int)crank * 256 + tid; x < sb_end - sb; x += 256 * CS) {
                double acc = 0.0;
                for (int r = 0; r < CS; ++r) acc += cluster.map_shared_rank(c_col, r)[x];
                ph2[sb + x] = acc;
            }
            cluster.sync();
        }
    }
    // ---- arg-max / counts / sums: warp -> CTA -> cluster, in rank order ---------------------------------
    for (int d = 16; d > 0; d >>= 1) {
        ArgMax o;
        o.ml = __shfl_down_sync(0xffffffffu, best.ml, d);
        o.h1 = __shfl_down_sync(0xffffffffu, best.h1, d);
        o.idx = __shfl_down_sync(0xffffffffu, best.idx, d);
        if (better(o, best)) best = o;
        cnt += __shfl_down_sync(0xffffffffu, cnt, d);
    }
    if (lane == 0) { s_best[warp] = best; s_cnt[warp] = cnt; s_sum[warp] = sum_all; s_path[warp] = sum_path; }
    __syncthreads();
    ArgMax top = s_best[0];
    int npoints = s_cnt[0];
    double tot_all = s_sum[0], tot_path = s_path[0];
    for (int w = 1; w < 8; ++w) {
        if (better(s_best[w], top)) top = s_best[w];
        npoints += s_cnt[w]; tot_all += s_sum[w]; tot_path += s_path[w];
    }
    if (CS > 1) {
        cg::cluster_group cluster = cg::this_cluster();
        if (tid == 0) { c_best = top; c_cnt = npoints; c_sum = tot_all; c_path = tot_path; }
        cluster.sync();
        if (crank == 0 && tid == 0) {
            top = ArgMax{-INFINITY, 0x7fffffff, 0x7fffffffffffffffLL}; npoints = 0; tot_all = 0.0; tot_path = 0.0;
            for (int r = 0; r < CS; ++r) {
                const ArgMax o = *cluster.map_shared_rank(&c_best, r);
                if (better(o, top)) top = o;
                npoints += *cluster.map_shared_rank(&c_cnt, r);
                tot_all += *cluster.map_shared_rank(&c_sum, r);
                tot_path += *cluster.map_shared_rank(&c_path, r);
            }
        }
        cluster.sync();
    }
    if (crank == 0 && tid == 0) {
        tredsw_grid_result r;
        r.max_ml = top_ml; r.sum_all = tot_all; r.sum_path = tot_path;
        r.arg_i1 = npoints ? (int)(top.idx / P.n_h2) : -1;
        r.arg_i2 = npoints ? (int)(top.idx % P.n_h2) : -1;
        r.n_points = npoints; r.pad = 0;
        g.res[pi] = r;
    }
    __syncthreads();
    }
}

// One warp per small surface (<= GRID_WARP_LIMIT points, including the empty ones of loci without evidence):
// the same reductions with shuffles only — a cohort step has thousands of surfaces of a few dozen points.
__global__ void __launch_bounds__(256) grid_reduce_warp_kernel(GridParams g, int nproblems) {
    const int lane = threadIdx.x & 31;
    const int pi = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (pi >= nproblems) return;
    const tredsw_grid_problem P = g.prob[pi];
    const long long total = P.n_h2 > 0 ? (long long)P.n_h1 * P.n_h2 : 0;
    if (total > GRID_WARP_LIMIT) return;
    const int32_t *h1s = g.ipool + P.off_h1, *h2s = g.ipool + P.off_h2;
    const double *surf = g.surface + P.off_surface;
    ArgMax best{-INFINITY, 0x7fffffff, 0x7fffffffffffffffLL};
    int cnt = 0;
    for (int t = lane; t < (int)total; t += 32) {
        const double ml = surf[t];
        if (ml == -INFINITY) continue;
        ++cnt;
        ArgMax c{ml, h1s[t / P.n_h2], (long long)t};
        if (better(c, best)) best = c;
    }
    for (int d = 16; d > 0; d >>= 1) {
        ArgMax o;
        o.ml = __shfl_down_sync(0xffffffffu, best.ml, d);
        o.h1 = __shfl_down_sync(0xffffffffu, best.h1, d);
        o.idx = __shfl_down_sync(0xffffffffu, best.idx, d);
        if (better(o, best)) best = o;
        cnt += __shfl_down_sync(0xffffffffu, cnt, d);
    }
    ArgMax top;
    top.ml = __shfl_sync(0xffffffffu, best.ml, 0);
    top.h1 = __shfl_sync(0xffffffffu, best.h1, 0);
    top.idx = __shfl_sync(0xffffffffu, best.idx, 0);
    const int npoints = __shfl_sync(0xffffffffu, cnt, 0);
    double *ph1 = g.marg + P.off_ph1, *ph2 = g.marg + P.off_ph2;
    double sum_all = 0.0, sum_path = 0.0;
    for (int i1 = 0; i1 < P.n_h1; ++i1) {
        const int h1 = h1s[i1];
        double acc = 0.0, accp = 0.0;
        for (int i2 = lane; i2 < P.n_h2; i2 += 32) {
            const double ml = surf[i1 * P.n_h2 + i2];
            if (ml == -INFINITY) continue;
            const double w = exp(ml - top.ml);
            acc += w;
            const int h2 = (P.ploidy == 1) ? h1 : h2s[i2];
            if (pathological(P, h1, h2)) accp += w;
        }
        for (int d = 16; d > 0; d >>= 1) {
            acc += __shfl_down_sync(0xffffffffu, acc, d);
            accp += __shfl_down_sync(0xffffffffu, accp, d);
        }
        if (lane == 0) { ph1[i1] = acc; sum_all += acc; sum_path += accp; }
    }
    for (int i2 = lane; i2 < P.n_h2; i2 += 32) {
        double acc = 0.0;
        for (int i1 = 0; i1 < P.n_h1; ++i1) {
            const double ml = surf[i1 * P.n_h2 + i2];
            if (ml == -INFINITY) continue;
            acc += exp(ml - top.ml);
        }
        ph2[i2] = acc;
    }
    if (lane == 0) {
        tredsw_grid_result r;
        r.max_ml = top.ml; r.sum_all = sum_all; r.sum_path = sum_path;
        r.arg_i1 = npoints ? (int)(top.idx / P.n_h2) : -1;
        r.arg_i2 = npoints ? (int)(top.idx % P.n_h2) : -1;
        r.n_points = npoints; r.pad = 0;
        g.res[pi] = r;
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
                         tredsw_grid_result *d_res, long long points_hint) {
    GridParams g{};
    g.prob = d_prob; g.ipool = d_ipool; g.dpool = d_dpool; g.surface = d_surface; g.marg = d_marg; g.res = d_res;
    g.small_value = exp(-10.0);
    g.really_small = exp(-100.0);
    g.log_small = log(g.small_value);
    int rc;
    const size_t pt_bytes = (((size_t)nproblems + 1) * sizeof(long long) + 15) & ~(size_t)15;
    if ((rc = ctx->d_tiles.ensure(2 * pt_bytes + (2 + 2 * (size_t)nproblems) * sizeof(int)))) return rc;
    long long *d_pt = ctx->d_tiles.as<long long>();
    long long *d_tl = reinterpret_cast<long long *>(ctx->d_tiles.as<unsigned char>() + pt_bytes);
    int *d_lists = reinterpret_cast<int *>(ctx->d_tiles.as<unsigned char>() + 2 * pt_bytes);
    ctx->mark(2);
    grid_tiles_kernel<<<1, 1024, 0, ctx->stream>>>(d_prob, nproblems, d_pt, d_tl, d_lists);
    {   // per-problem constants and far-region tables of the medium / large surfaces
        const size_t far_bytes = (((size_t)nproblems * sizeof(FarInfo)) + 255) & ~(size_t)255;
        // table arena, in doubles: 64 MB serve ~250 long-expansion surfaces (a surface that finds it full is
        // evaluated point by point — slower, same result); a cohort searched with --fullsearch has one large
        // surface per problem, ~80 doubles of tables per candidate allele each
        long long cap = 8LL << 20;
        if (points_hint > 1024) {
            const long long side = (long long)ceil(sqrt((double)points_hint));
            const long long want = (long long)nproblems * 80 * side;
            if (want > cap) cap = want < (1LL << 30) ? want : (1LL << 30);
        }
        if (ctx->d_ftab.ensure(far_bytes + 256 + (size_t)cap * sizeof(double)) != TREDSW_OK) {
            cudaGetLastError();                                            // not enough memory for the big arena
            cap = 8LL << 20;
            if ((rc = ctx->d_ftab.ensure(far_bytes + 256 + (size_t)cap * sizeof(double)))) return rc;
        }
        g.far = ctx->d_ftab.as<FarInfo>();
        g.fcursor = reinterpret_cast<unsigned long long *>(ctx->d_ftab.as<unsigned char>() + far_bytes);
        g.ftab = reinterpret_cast<double *>(ctx->d_ftab.as<unsigned char>() + far_bytes + 256);
        g.ftab_cap = cap;
        CUDA_TRY(cudaMemsetAsync(g.far, 0, far_bytes + 256, ctx->stream));
        const int nbf = nproblems < ctx->sm_count * 4 ? nproblems : ctx->sm_count * 4;
        grid_setup_kernel<<<nbf, 256, 0, ctx->stream>>>(g, d_lists, nproblems);
        const long long want = (long long)nproblems * FILL_SPLIT;
        const int nfill = want < (long long)ctx->sm_count * 8 ? (int)want : ctx->sm_count * 8;
        grid_fill_kernel<<<nfill, 256, 0, ctx->stream>>>(g, d_lists, nproblems);
        ctx->launches += 2;
    }
    grid_surface_points_kernel<<<ctx->sm_count * 4, GRID_TILE, 0, ctx->stream>>>(g, nproblems, d_pt);
    grid_surface_tiles_kernel<<<ctx->sm_count * 8, GRID_TILE, 0, ctx->stream>>>(g, nproblems, d_tl);
    CUDA_TRY(cudaGetLastError());
    ctx->launches += 3;
    // reductions by class: a warp per small surface, a CTA per medium one, a cluster of 8 CTAs per large one
    grid_reduce_warp_kernel<<<(nproblems + 7) / 8, 256, 0, ctx->stream>>>(g, nproblems);
    const int nb_block = nproblems < ctx->sm_count * 4 ? nproblems : ctx->sm_count * 4;
    grid_reduce_kernel<1><<<nb_block, 256, 0, ctx->stream>>>(g, d_lists + 2, d_lists);
    CUDA_TRY(cudaGetLastError());
    {
        const int nclusters = nproblems < ctx->sm_count ? nproblems : ctx->sm_count;   // persistent over the class list
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)nclusters * GRID_CLUSTER); cfg.blockDim = dim3(256); cfg.stream = ctx->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = GRID_CLUSTER; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        const int *cl_list = d_lists + 2 + nproblems, *cl_n = d_lists + 1;
        CUDA_TRY(cudaLaunchKernelEx(&cfg, grid_reduce_kernel<GRID_CLUSTER>, g, cl_list, cl_n));
    }
    ctx->mark(3);
    ctx->launches += 3;
    return TREDSW_OK;
}

extern "C" int tredsw_likelihood_grid(tredsw_ctx *ctx, const tredsw_grid_problem *problems, int32_t nproblems,
                                      const int32_t *ipool, int64_t n_ipool, const double *dpool,
                                      int64_t n_dpool, double *surface, int64_t n_surface, double *marg,
                                      int64_t n_marg, tredsw_grid_result *results, uint32_t flags) {
    if (!ctx) { tredsw_set_error("null context"); return TREDSW_ERR_ARG; }
    if (nproblems < 0 || !problems || !results) { tredsw_set_error("bad arguments"); return TREDSW_ERR_ARG; }
    if (nproblems == 0) return TREDSW_OK;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    GridParams g{};
    int rc;
    long long max_points = 0;
    if (!dev_ptrs(flags)) {
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
    if (dev_ptrs(flags)) { g.surface = surface; g.marg = marg; g.res = results; }
    else {
        if ((rc = ctx->d_surface.ensure((size_t)(n_surface > 0 ? n_surface : 1) * sizeof(double)))) return rc;
        if ((rc = ctx->d_marg.ensure((size_t)(n_marg > 0 ? n_marg : 1) * sizeof(double)))) return rc;
        if ((rc = ctx->d_res.ensure((size_t)nproblems * sizeof(tredsw_grid_result)))) return rc;
        g.surface = ctx->d_surface.as<double>(); g.marg = ctx->d_marg.as<double>();
        g.res = ctx->d_res.as<tredsw_grid_result>();
    }
    if ((rc = tredsw_internal_grid(ctx, g.prob, nproblems, g.ipool, g.dpool, g.surface, g.marg, g.res, max_points))) return rc;
    if (!dev_ptrs(flags)) {
        if (surface && n_surface > 0)
            CUDA_TRY(cudaMemcpyAsync(surface, g.surface, (size_t)n_surface * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (marg && n_marg > 0)
            CUDA_TRY(cudaMemcpyAsync(marg, g.marg, (size_t)n_marg * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(results, g.res, (size_t)nproblems * sizeof(tredsw_grid_result), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
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

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
    // far-region tables of the large surfaces (grid_far_kernel)
    struct FarInfo *far;            // [nproblems]
    double *ftab;                   // arena of per-problem tables
    long long ftab_cap;
    unsigned long long *fcursor;    // arena cursor
};

// In the FAR region of a large surface — both alleles beyond every observed key, the partial clamp and the
// read length, and the longer allele shifted past the KDE support — the four terms collapse:
//   spanning + partial = one constant c12;  repeat-only = a function of h1 + h2 only;  paired-end = a function
//   of h1 only.
// The candidate lists end in arithmetic progressions of step `period` there, so for list indices i1 >= fa1,
// i2 >= fa2 the point is   ml = ((c12) + rept[(i1-fa1) + (i2-fa2)]) + pe_row[i1-fa1]   — the same three
// additions in the same order as the general evaluation, on operands produced by the same code, hence the
// same bits — two table reads instead of tens of logarithms.  On a --fullsearch / long-expansion grid
// (10^5 - 10^6 points) ~90 % of the points are of this kind.
struct FarInfo {
    long long off;      // ftab: pe_row[n_h1 - fa1] then rept[(n_h1 - fa1) + (n_h2 - fa2) - 1]
    double c12;
    int fa1, fa2, ok, pad;
};

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
    int far_ok, fa1, fa2;  // far region of this problem (FarInfo)
    double c12;
    const double *pe_row, *rept;
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
    const double xl = (P.n_rept == 0) ? 0.0 : kk * log(mu);
    const double pk = xl - lgamma_k1 - mu;
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
        if (t > GRID_WARP_LIMIT) ts += (t + GRID_TILE - 1) / GRID_TILE; else s += t;
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
        if (t > GRID_WARP_LIMIT) trun += (t + GRID_TILE - 1) / GRID_TILE; else run += t;
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
    T.far_ok = 0; T.fa1 = T.fa2 = 0; T.c12 = 0.0; T.pe_row = T.rept = nullptr;
    return T;
}

// One block per medium / large surface (class lists of grid_tiles_kernel): decide the far region and fill its
// tables.
__global__ void __launch_bounds__(256) grid_far_kernel(GridParams g, const int *lists, int nproblems) {
    __shared__ FarInfo s_f;
    __shared__ TileShared s_t;
    const int nb = lists[0], nc = lists[1];
    for (int k = blockIdx.x; k < nb + nc; k += gridDim.x) {
        const int pi = k < nb ? lists[2 + k] : lists[2 + nproblems + (k - nb)];
        const tredsw_grid_problem &P = g.prob[pi];
        const int32_t *h1s = g.ipool + P.off_h1, *h2s = g.ipool + P.off_h2;
        if (threadIdx.x == 0) {
            FarInfo f; f.off = 0; f.c12 = 0.0; f.fa1 = f.fa2 = 0; f.ok = 0; f.pad = 0;
            TileShared T = tile_shared_of(g, P);
            if (P.ploidy != 1 && P.n_h1 > 0 && P.n_h2 > 1) {
                const int32_t *skey = g.ipool + P.off_span;
                int ks = -1000000;
                for (int i = 0; i < P.n_span; ++i) ks = max(ks, skey[i]);
                const int t1 = P.readlen - 9;
                // h >= H1: no spanning key within 18, partial clamp and mixing weight saturated, h > readlen
                const int H1 = max(max(ks + DEV + 1, P.max_partial), max(t1, P.readlen + 1));
                // h >= H2: every pair length >= MINPE is shifted past the support
                long long H2 = H1;
                if (P.run_pe && T.tmin != 0x7fffffff) H2 = max((long long)H1, (long long)P.pe_ref + SPAN - T.tmin);
                auto tail = [&](const int32_t *hs, int n, long long H) {
                    int i = n - 1;
                    if (hs[i] < H) return n;
                    while (i > 0 && hs[i] - hs[i - 1] == P.period && hs[i - 1] >= H) --i;
                    return i;
                };
                f.fa1 = tail(h1s, P.n_h1, H1);
                f.fa2 = tail(h2s, P.n_h2, H2);
                const long long r1 = P.n_h1 - f.fa1, r2 = P.n_h2 - f.fa2;
                if (r1 > 0 && r2 > 0 && r1 * r2 >= 4096) {
                    const unsigned long long need = (unsigned long long)(r1 + r1 + r2);
                    const unsigned long long off = atomicAdd(g.fcursor, need);
                    if ((long long)(off + need) <= g.ftab_cap) {
                        f.off = (long long)off; f.ok = 1;
                        const int h1 = h1s[f.fa1], h2 = h2s[f.fa2];
                        f.c12 = __dadd_rn(ml_span_term(P, g, h1, h2), ml_part_term(P, g, h1, h2, T.sig_mp));
                    }
                }
            }
            s_f = f; s_t = T;
            g.far[pi] = f;
        }
        __syncthreads();
        const FarInfo f = s_f;
        if (f.ok) {
            const TileShared T = s_t;
            const int r1 = P.n_h1 - f.fa1, r2 = P.n_h2 - f.fa2;
            double *pe_row = g.ftab + f.off, *rept = pe_row + r1;
            const int hfar = h2s[f.fa2];                       // any allele past the support
            for (int r = threadIdx.x; r < r1; r += blockDim.x) pe_row[r] = ml_pe_term(P, g, h1s[f.fa1 + r], hfar, T.tmin);
            const int base = h1s[f.fa1] + h2s[f.fa2] - 2 * P.readlen;   // d1 + d2 at (fa1, fa2); both alleles > readlen
            for (int r = threadIdx.x; r < r1 + r2 - 1; r += blockDim.x) rept[r] = ml_rept_term(P, base + r * P.period, T.lgamma_k1);
        }
        __syncthreads();
    }
}

// Large and medium surfaces: persistent over their tiles; the per-problem values are computed once per tile.
__global__ void __launch_bounds__(GRID_TILE) grid_surface_tiles_kernel(GridParams g, int nproblems, const long long *tile_start) {
    const long long ntiles = tile_start[nproblems];
    __shared__ int s_pi;
    __shared__ TileShared s_tile;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        if (threadIdx.x == 0) {                        // tile -> problem: last p with tile_start[p] <= tile
            int lo = 0, hi = nproblems;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (tile_start[mid] <= tile) lo = mid; else hi = mid;
            }
            s_pi = lo;
            TileShared T0 = tile_shared_of(g, g.prob[lo]);
            const FarInfo f = g.far[lo];
            if (f.ok) {
                T0.far_ok = 1; T0.fa1 = f.fa1; T0.fa2 = f.fa2; T0.c12 = f.c12;
                T0.pe_row = g.ftab + f.off; T0.rept = T0.pe_row + (g.prob[lo].n_h1 - f.fa1);
            }
            s_tile = T0;
        }
        __syncthreads();
        const int pi = s_pi;
        const TileShared T = s_tile;
        __syncthreads();
        const tredsw_grid_problem &P = g.prob[pi];
        const long long total = (long long)P.n_h1 * P.n_h2;
        const long long t = (tile - tile_start[pi]) * GRID_TILE + threadIdx.x;
        if (t < total) {
            const int32_t *h1s = g.ipool + P.off_h1, *h2s = g.ipool + P.off_h2;
            int i1, i2;
            if (total < 0x7fffffffLL) { i1 = (int)((unsigned)t / (unsigned)P.n_h2); i2 = (int)((unsigned)t - (unsigned)i1 * (unsigned)P.n_h2); }
            else { i1 = (int)(t / P.n_h2); i2 = (int)(t % P.n_h2); }
            const int h1 = h1s[i1];
            const int h2 = (P.ploidy == 1) ? h1 : h2s[i2];
            double ml = -INFINITY;
            if (h1 <= h2) {
                if (T.far_ok && i1 >= T.fa1 && i2 >= T.fa2)
                    ml = __dadd_rn(__dadd_rn(T.c12, T.rept[(i1 - T.fa1) + (i2 - T.fa2)]), T.pe_row[i1 - T.fa1]);
                else
                    ml = point_ml(P, g, h1, h2, T);
            }
            g.surface[P.off_surface + t] = ml;
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

__device__ __forceinline__ bool pathological(const tredsw_grid_problem &P, int h1, int h2) {
    const int lo = min(h1, h2) / P.period, hi = max(h1, h2) / P.period;
    if (P.expansion) return P.recessive ? (lo >= P.cutoff_risk) : (hi >= P.cutoff_risk);
    return P.recessive ? (hi <= P.cutoff_risk) : (lo <= P.cutoff_risk);
}

// Reductions of one problem's surface (models.py:277-302, 342-368): max / arg-max with key (ml, -h1, order)
// (Q10), the exp-normalised marginals P_h1 / P_h2 and the PP sums.  CS = 1: one CTA per problem (medium
// surfaces).  CS = 8: a thread-block CLUSTER of 8 CTAs per problem for the large surfaces (extended ranges,
// --fullsearch: 10^4 - 10^6 points): rows / columns are dealt to the CTAs of the cluster, the per-CTA partial
// results meet in distributed shared memory in rank order, so the result is deterministic.  Both walk the
// index list of their class (grid_tiles_kernel); small surfaces take grid_reduce_warp_kernel.
template <int CS>
__global__ void __launch_bounds__(256, 4) grid_reduce_kernel(GridParams g, const int *list, const int *nlist) {
    namespace cg = cooperative_groups;
    __shared__ ArgMax s_best[256];
    __shared__ double s_sum[256], s_path[256];
    __shared__ int s_cnt[256];
    __shared__ ArgMax c_best;          // this CTA's partial results, read by the other CTAs of the cluster
    __shared__ int c_cnt;
    __shared__ double c_sum, c_path;
    unsigned crank = 0;
    if (CS > 1) crank = cg::this_cluster().block_rank();
    const int n = *nlist;
    // persistent over the problems of this class (CTAs of a cluster walk the list together)
    for (int k = blockIdx.x / CS; k < n; k += (CS > 1 ? n : (int)gridDim.x)) {       // (cluster variant: one problem per cluster)
    const int pi = list[k];
    const tredsw_grid_problem P = g.prob[pi];
    const int32_t *h1s = g.ipool + P.off_h1, *h2s = g.ipool + P.off_h2;
    const double *surf = g.surface + P.off_surface;
    const int tid = threadIdx.x;
    // ---- max / arg-max / number of evaluated points: rows crank, crank + CS, ... ----------------------
    ArgMax best{-INFINITY, 0x7fffffff, 0x7fffffffffffffffLL};
    int cnt = 0;
    for (int i1 = (int)crank; i1 < P.n_h1; i1 += CS) {
        const int h1 = h1s[i1];
        const long long row = (long long)i1 * P.n_h2;
        for (int i2 = tid; i2 < P.n_h2; i2 += 256) {
            const double ml = surf[row + i2];
            if (ml == -INFINITY) continue;   // not evaluated (h1 > h2)
            ++cnt;
            ArgMax c{ml, h1, row + i2};
            if (better(c, best)) best = c;
        }
    }
    s_best[tid] = best; s_cnt[tid] = cnt;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1) {
        if (tid < d) {
            if (better(s_best[tid + d], s_best[tid])) s_best[tid] = s_best[tid + d];
            s_cnt[tid] += s_cnt[tid + d];
        }
        __syncthreads();
    }
    ArgMax top = s_best[0];
    int npoints = s_cnt[0];
    if (CS > 1) {
        cg::cluster_group cluster = cg::this_cluster();
        if (tid == 0) { c_best = top; c_cnt = npoints; }
        cluster.sync();
        top = ArgMax{-INFINITY, 0x7fffffff, 0x7fffffffffffffffLL}; npoints = 0;
        for (int r = 0; r < CS; ++r) {
            const ArgMax o = *cluster.map_shared_rank(&c_best, r);
            if (better(o, top)) top = o;
            npoints += *cluster.map_shared_rank(&c_cnt, r);
        }
        cluster.sync();
    }
    __syncthreads();
    // ---- marginals: P_h1[i1] = sum_i2 w, P_h2[i2] = sum_i1 w, w = exp(ml - max) ------------------------
    double *ph1 = g.marg + P.off_ph1, *ph2 = g.marg + P.off_ph2;
    double sum_all = 0.0, sum_path = 0.0;
    const int warp = tid >> 5, lane = tid & 31, nwarps = 8;
    for (int i1 = (int)crank * nwarps + warp; i1 < P.n_h1; i1 += nwarps * CS) {
        const int h1 = h1s[i1];
        double acc = 0.0, accp = 0.0;
        for (int i2 = lane; i2 < P.n_h2; i2 += 32) {
            const double ml = surf[(long long)i1 * P.n_h2 + i2];
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
    for (int i2 = (int)crank * 256 + tid; i2 < P.n_h2; i2 += 256 * CS) {
        double acc = 0.0;
        for (int i1 = 0; i1 < P.n_h1; ++i1) {
            const double ml = surf[(long long)i1 * P.n_h2 + i2];
            if (ml == -INFINITY) continue;
            acc += exp(ml - top.ml);
        }
        ph2[i2] = acc;
    }
    s_sum[tid] = sum_all; s_path[tid] = sum_path;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1) {
        if (tid < d) { s_sum[tid] += s_sum[tid + d]; s_path[tid] += s_path[tid + d]; }
        __syncthreads();
    }
    double tot_all = s_sum[0], tot_path = s_path[0];
    if (CS > 1) {
        cg::cluster_group cluster = cg::this_cluster();
        if (tid == 0) { c_sum = tot_all; c_path = tot_path; }
        cluster.sync();
        tot_all = 0.0; tot_path = 0.0;
        if (crank == 0 && tid == 0)
            for (int r = 0; r < CS; ++r) { tot_all += *cluster.map_shared_rank(&c_sum, r); tot_path += *cluster.map_shared_rank(&c_path, r); }
        cluster.sync();
    }
    if (crank == 0 && tid == 0) {
        tredsw_grid_result r;
        r.max_ml = top.ml; r.sum_all = tot_all; r.sum_path = tot_path;
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
    (void)points_hint;
    int rc;
    const size_t pt_bytes = (((size_t)nproblems + 1) * sizeof(long long) + 15) & ~(size_t)15;
    if ((rc = ctx->d_tiles.ensure(2 * pt_bytes + (2 + 2 * (size_t)nproblems) * sizeof(int)))) return rc;
    long long *d_pt = ctx->d_tiles.as<long long>();
    long long *d_tl = reinterpret_cast<long long *>(ctx->d_tiles.as<unsigned char>() + pt_bytes);
    int *d_lists = reinterpret_cast<int *>(ctx->d_tiles.as<unsigned char>() + 2 * pt_bytes);
    ctx->mark(2);
    grid_tiles_kernel<<<1, 1024, 0, ctx->stream>>>(d_prob, nproblems, d_pt, d_tl, d_lists);
    {   // far-region tables of the medium / large surfaces
        const size_t far_bytes = (((size_t)nproblems * sizeof(FarInfo)) + 255) & ~(size_t)255;
        const long long cap = 8LL << 20;                                   // doubles (64 MB): tables of ~2000 large surfaces
        if ((rc = ctx->d_ftab.ensure(far_bytes + 256 + (size_t)cap * sizeof(double)))) return rc;
        g.far = ctx->d_ftab.as<FarInfo>();
        g.fcursor = reinterpret_cast<unsigned long long *>(ctx->d_ftab.as<unsigned char>() + far_bytes);
        g.ftab = reinterpret_cast<double *>(ctx->d_ftab.as<unsigned char>() + far_bytes + 256);
        g.ftab_cap = cap;
        CUDA_TRY(cudaMemsetAsync(g.far, 0, far_bytes + 256, ctx->stream));
        const int nbf = nproblems < ctx->sm_count * 4 ? nproblems : ctx->sm_count * 4;
        grid_far_kernel<<<nbf, 256, 0, ctx->stream>>>(g, d_lists, nproblems);
        ctx->launches += 1;
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
        const int nclusters = nproblems;      // one cluster per problem; clusters beyond the class list exit at once
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

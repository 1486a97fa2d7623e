// kde.cuh — Gaussian KDE of paired-end lengths on the grid 0..999, one 256-thread block per problem.
//
// Replaces PEMaxLikModel.__init__ (tredparse/models.py:428-435): scipy.stats.gaussian_kde with Scott's
// factor n^(-1/5) — covariance = var(ddof=1) * factor^2, pdf[x] ~ sum_i exp(-((x - x_i)/sd)^2 / 2) — then
// pdf / pdf.sum() (the normalisation constant cancels).  Lengths are integers, so the sum runs over a
// histogram of the distinct values (<= 2048 terms per grid point instead of one per pair).
// The Gaussian is tabulated per problem over the integer distances (the arguments differ from scipy's
// (x_i - x) / sd by one rounding, ~1e-16 relative — the parity bar on the likelihood is 1e-9).
// Lengths outside [-1024, 1024) are clamped into the histogram (the reference keeps only tlen < 1000,
// bam_parser.py:356-357; large negative lengths do not occur for properly oriented pairs).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int KDE_SPAN = 1000;
constexpr int KDE_OFF = 1024;

// Block size: KDE_THREADS threads, KDE_SPAN / KDE_THREADS grid points each.  (256 threads, not one thread
// per grid point: a 1024-thread block cannot become resident next to the persistent Smith-Waterman CTAs of a
// concurrently running call — registers — and would stall its whole call behind them.)
constexpr int KDE_THREADS = 256;
constexpr int KDE_PER_THREAD = (KDE_SPAN + KDE_THREADS - 1) / KDE_THREADS;

__device__ __forceinline__ double kde_block_sum(double *red, double v) {
    const int tid = threadIdx.x;
    red[tid] = v;
    __syncthreads();
    for (int d = KDE_THREADS / 2; d > 0; d >>= 1) { if (tid < d) red[tid] += red[tid + d]; __syncthreads(); }
    const double r = red[0];
    __syncthreads();
    return r;
}

__device__ __forceinline__ void kde_block(const int32_t *x, int n, double *out) {
    __shared__ int hist[2 * KDE_OFF];
    __shared__ double red[KDE_THREADS];
    const int tid = threadIdx.x;
    for (int i = tid; i < 2 * KDE_OFF; i += KDE_THREADS) hist[i] = 0;
    __syncthreads();
    double s = 0.0;
    for (int i = tid; i < n; i += KDE_THREADS) {
        int v = x[i];
        s += (double)v;
        v = max(-KDE_OFF, min(KDE_OFF - 1, v));
        atomicAdd(&hist[v + KDE_OFF], 1);
    }
    const double mean = n > 0 ? kde_block_sum(red, s) / (double)n : (kde_block_sum(red, s), 0.0);
    double ss = 0.0;
    for (int i = tid; i < n; i += KDE_THREADS) { double d = (double)x[i] - mean; ss += d * d; }
    const double tot = kde_block_sum(red, ss);
    const double var = n > 1 ? tot / (double)(n - 1) : 0.0;
    const double factor = pow((double)n, -1.0 / 5.0);
    const double sd = sqrt(var * factor * factor);
    // compact the occupied histogram bins in ascending order (deterministic: the sums below run in the same
    // order as a plain scan over all bins); typically a few hundred of the 2048 bins are occupied
    // (the compacted counts overwrite the histogram — every thread holds its bins in registers first — so the
    // block needs 31 KB of shared memory and fits next to the Smith-Waterman CTAs of a concurrent call)
    __shared__ short cbin[2 * KDE_OFF];
    int *ccnt = hist;
    __shared__ int scan[KDE_THREADS];
    constexpr int PER = 2 * KDE_OFF / KDE_THREADS;
    int mine_n = 0, mine_c[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) { mine_c[k] = hist[tid * PER + k]; mine_n += mine_c[k] != 0; }
    scan[tid] = mine_n;
    __syncthreads();
    for (int d = 1; d < KDE_THREADS; d <<= 1) {
        const int v = tid >= d ? scan[tid - d] : 0;
        __syncthreads();
        scan[tid] += v;
        __syncthreads();
    }
    const int nbins = scan[KDE_THREADS - 1];
    int w = scan[tid] - mine_n;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int c = mine_c[k];
        if (c != 0) { cbin[w] = (short)(tid * PER + k - KDE_OFF); ccnt[w] = c; ++w; }
    }
    __syncthreads();
    // The kernel value depends on the integer distance |bin - x| only: tabulate exp(-(d / sd)^2 / 2) once
    // (2048 exps per problem), then every grid point is a dot product of the occupied bins with the table.
    __shared__ double gtab[2 * KDE_OFF];
    for (int d = tid; d < 2 * KDE_OFF; d += KDE_THREADS) {
        const double r = (double)d / sd;
        gtab[d] = exp(-(r * r) / 2.0);
    }
    __syncthreads();
    double val[KDE_PER_THREAD];
    double mine = 0.0;
#pragma unroll
    for (int k = 0; k < KDE_PER_THREAD; ++k) {
        const int p = tid + k * KDE_THREADS;
        val[k] = 0.0;
        if (p < KDE_SPAN && n > 0) {
            double acc = 0.0;
            for (int b = 0; b < nbins; ++b) {
                const int d = abs((int)cbin[b] - p);            // <= 1024 + 999 < 2 * KDE_OFF
                acc += (double)ccnt[b] * gtab[d];
            }
            val[k] = acc;
        }
        mine += val[k];
    }
    const double total = kde_block_sum(red, mine);
#pragma unroll
    for (int k = 0; k < KDE_PER_THREAD; ++k) {
        const int p = tid + k * KDE_THREADS;
        if (p < KDE_SPAN) out[p] = val[k] / total;
    }
    __syncthreads();
}

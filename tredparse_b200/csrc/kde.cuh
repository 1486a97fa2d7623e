// kde.cuh — Gaussian KDE of paired-end lengths on the grid 0..999, one 1024-thread block per problem.
//
// Replaces PEMaxLikModel.__init__ (tredparse/models.py:428-435): scipy.stats.gaussian_kde with Scott's
// factor n^(-1/5) — covariance = var(ddof=1) * factor^2, pdf[x] ~ sum_i exp(-((x - x_i)/sd)^2 / 2) — then
// pdf / pdf.sum() (the normalisation constant cancels).  Lengths are integers, so the sum runs over a
// histogram of the distinct values (<= 2048 terms per grid point instead of one per pair).
// Lengths outside [-1024, 1024) are clamped into the histogram (the reference keeps only tlen < 1000,
// bam_parser.py:356-357; large negative lengths do not occur for properly oriented pairs).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int KDE_SPAN = 1000;
constexpr int KDE_OFF = 1024;

__device__ __forceinline__ void kde_block(const int32_t *x, int n, double *out) {
    __shared__ int hist[2 * KDE_OFF];
    __shared__ double red[1024];
    __shared__ double s_mean, s_sd;
    const int tid = threadIdx.x;
    for (int i = tid; i < 2 * KDE_OFF; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    double s = 0.0;
    for (int i = tid; i < n; i += blockDim.x) {
        int v = x[i];
        s += (double)v;
        v = max(-KDE_OFF, min(KDE_OFF - 1, v));
        atomicAdd(&hist[v + KDE_OFF], 1);
    }
    red[tid] = s;
    __syncthreads();
    for (int d = 512; d > 0; d >>= 1) { if (tid < d) red[tid] += red[tid + d]; __syncthreads(); }
    if (tid == 0) s_mean = n > 0 ? red[0] / (double)n : 0.0;
    __syncthreads();
    double ss = 0.0;
    for (int i = tid; i < n; i += blockDim.x) { double d = (double)x[i] - s_mean; ss += d * d; }
    red[tid] = ss;
    __syncthreads();
    for (int d = 512; d > 0; d >>= 1) { if (tid < d) red[tid] += red[tid + d]; __syncthreads(); }
    if (tid == 0) {
        const double var = n > 1 ? red[0] / (double)(n - 1) : 0.0;
        const double factor = pow((double)n, -1.0 / 5.0);
        s_sd = sqrt(var * factor * factor);
    }
    __syncthreads();
    double val = 0.0;
    if (tid < KDE_SPAN && n > 0) {
        const double xs = (double)tid / s_sd;
        for (int b = 0; b < 2 * KDE_OFF; ++b) {
            const int c = hist[b];
            if (c == 0) continue;
            const double r = (double)(b - KDE_OFF) / s_sd - xs;
            val += (double)c * exp(-(r * r) / 2.0);
        }
    }
    red[tid] = tid < KDE_SPAN ? val : 0.0;
    __syncthreads();
    for (int d = 512; d > 0; d >>= 1) { if (tid < d) red[tid] += red[tid + d]; __syncthreads(); }
    if (tid < KDE_SPAN) out[tid] = val / red[0];
    __syncthreads();
}

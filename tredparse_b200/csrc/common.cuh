// common.cuh — context, error plumbing and device workspace shared by the libtredsw translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <vector>
#include <mutex>

#include "../../include/tredsw.h"

#define TREDSW_VERSION 100

void tredsw_set_error(const char *fmt, ...);

#define CUDA_TRY(expr)                                                                        \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            tredsw_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                             __LINE__);                                                       \
            return TREDSW_ERR_CUDA;                                                           \
        }                                                                                     \
    } while (0)

// A grow-only device buffer.
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return TREDSW_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        CUDA_TRY(cudaMalloc(&p, want));
        cap = want;
        return TREDSW_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() { return reinterpret_cast<T *>(p); }
};

struct tredsw_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 0;
    size_t smem_optin = 0;
    // staging buffers (host-pointer mode)
    DevBuf d_q, d_qoff, d_t, d_toff, d_qidx, d_tidx, d_out, d_cigar, d_scratch, d_misc, d_fam,
        d_rfam, d_stats, d_work, d_prob, d_ipool, d_dpool, d_surface, d_marg, d_res, d_counter, d_tiles, d_pk, d_pe16, d_ftab;
    std::mutex mu;
    long long launches = 0;
    bool timing = false;
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool ev_valid[8] = {false, false, false, false, false, false, false, false};
    void mark(int i) {            // record event i on the stream when timing is enabled
        if (!timing) return;
        if (!ev[i]) cudaEventCreate(&ev[i]);
        cudaEventRecord(ev[i], stream);
        ev_valid[i] = true;
    }
    ~tredsw_ctx();
};

static inline bool dev_ptrs(uint32_t flags) { return (flags & TREDSW_DEVICE_PTRS) != 0; }

// Copy helper: returns the device pointer to use for `host_or_dev`.
template <class T>
static inline int stage_in(tredsw_ctx *ctx, DevBuf &buf, const T *src, size_t n, uint32_t flags,
                           const T **out) {
    if (dev_ptrs(flags) || n == 0) { *out = src; if (n == 0 && !dev_ptrs(flags)) { int rc = buf.ensure(16); if (rc) return rc; *out = buf.as<T>(); } return TREDSW_OK; }
    int rc = buf.ensure(n * sizeof(T));
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(buf.p, src, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    *out = buf.as<T>();
    return TREDSW_OK;
}

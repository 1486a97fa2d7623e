// sw_pairs.cu — generic batch of independent (query, template) alignments, one thread per pair.
//
// Replaces a batch of ssw_init -> ssw_align(flag=1) -> init_destroy -> align_destroy calls
// (src/ssw_wrap.py:186-224 over src/ssw.c:751-871).  Outputs per pair: score, ref_begin, ref_end,
// query_begin, query_end (bit-exact) and — on request — score2 / ref_end2
// (src/ssw.c:315-340, 519-542, including the ghost padding rows of the striped layout) and the CIGAR
// (banded_sw restatement, src/ssw.c:549-736).
//
// This is the compatibility / verification path (any sequences, any lengths); the throughput path
// for tredparse's workload is the family kernel in sw_family.cu.
#include "common.cuh"
#include "sw_sweep.cuh"

namespace {

constexpr int PAIRS_W = 16;
constexpr int PAIRS_THREADS = 64;

struct PairsParams {
    const int8_t *qbuf; const int64_t *qoff;
    const int8_t *tbuf; const int64_t *toff;
    const int32_t *qidx; const int32_t *tidx;
    int64_t npairs;
    int go, ge, bias;
    uint32_t flags;
    int32_t *out;
    uint32_t *bnd;        // [max_rows][nthreads]
    int32_t *colmax;      // [max_n][nthreads] (score2 only)
    int max_rows, max_n;
    uint32_t *cigar_out; int cigar_cap;
    int8_t *dirs; int64_t dirs_per_thread; int32_t *hbuf; int hbuf_per_thread;
};

__constant__ int8_t c_mat25[25];

// banded_sw restatement (src/ssw.c:549-736), one thread, scratch in global memory.
// h_b/e_b/h_c: 3 arrays of `width+1` ints; dirs: width_d*readLen*3 bytes. Returns cigar length or -1.
__device__ int banded_cigar(const int8_t *ref, const int8_t *read, int refLen, int readLen, int score,
                            int go, int ge, int32_t *hbuf, int hbuf_cap, int8_t *dirs, int64_t dirs_cap,
                            uint32_t *out, int cap) {
    int band_width = abs(refLen - readLen) + 1;
    int width = 0, width_d = 0, mx = 0;
    int guard = 0;
    do {
        width = band_width * 2 + 3; width_d = band_width * 2 + 1;
        if ((width + 1) * 3 > hbuf_cap || (int64_t)width_d * readLen * 3 > dirs_cap || ++guard > 32) return -1;
        int32_t *h_b = hbuf, *e_b = hbuf + (width + 1), *h_c = hbuf + 2 * (width + 1);
        mx = 0;
        for (int j = 1; j < width - 1; ++j) h_b[j] = 0;
        for (int i = 0; i < readLen; ++i) {
            int beg = 0, end = refLen - 1, u = 0, j;
            j = i - band_width; beg = beg > j ? beg : j;
            j = i + band_width; end = end < j ? end : j;
            int edge = end + 1 < width - 1 ? end + 1 : width - 1;
            int f = 0;
            h_b[0] = e_b[0] = h_b[edge] = e_b[edge] = h_c[0] = 0;
            int8_t *dl = dirs + (int64_t)width_d * i * 3;
            const int x0 = (i - band_width) > 0 ? (i - band_width) : 0;          // set_u / set_d offsets
            const int x1 = (i - 1 - band_width) > 0 ? (i - 1 - band_width) : 0;
            for (j = beg; j <= end; ++j) {
                u = j - x0 + 1;
                int e = j - x1 + 1;
                int b = j - 1 - x0 + 1;
                int d = j - 1 - x1 + 1;
                int de = (j - x0) * 3, df = de + 1, dh = de + 2;
                int temp1 = i == 0 ? -go : h_b[e] - go;
                int temp2 = i == 0 ? -ge : e_b[e] - ge;
                e_b[u] = temp1 > temp2 ? temp1 : temp2;
                dl[de] = temp1 > temp2 ? 3 : 2;
                temp1 = h_c[b] - go;
                temp2 = f - ge;
                f = temp1 > temp2 ? temp1 : temp2;
                dl[df] = temp1 > temp2 ? 5 : 4;
                int e1 = e_b[u] > 0 ? e_b[u] : 0;
                int f1 = f > 0 ? f : 0;
                temp1 = e1 > f1 ? e1 : f1;
                temp2 = h_b[d] + c_mat25[(int)ref[j] * 5 + (int)read[i]];
                h_c[u] = temp1 > temp2 ? temp1 : temp2;
                if (h_c[u] > mx) mx = h_c[u];
                if (temp1 <= temp2) dl[dh] = 1;
                else dl[dh] = e1 > f1 ? dl[de] : dl[df];
            }
            for (j = 1; j <= u; ++j) h_b[j] = h_c[j];
        }
        band_width *= 2;
    } while (mx < score);
    band_width /= 2;
    // trace back, writing the run-length ops backwards into out[cap-1 ...] then shifting to the front
    int i = readLen - 1, j = refLen - 1, e = 0, l = 0, state = 2;
    int op = 0, prev_op = 0;          // 0=M 1=I 2=D
    int8_t *dl = dirs + (int64_t)width_d * (readLen - 1) * 3;
    while (i > 0) {
        int x = (i - band_width) > 0 ? (i - band_width) : 0;
        int idx = (j - x) * 3 + state;
        switch (dl[idx]) {
            case 1: --i; --j; state = 2; dl -= width_d * 3; op = 0; break;
            case 2: --i; state = 0; dl -= width_d * 3; op = 1; break;
            case 3: --i; state = 2; dl -= width_d * 3; op = 1; break;
            case 4: --j; state = 1; op = 2; break;
            case 5: --j; state = 2; op = 2; break;
            default: return -1;
        }
        if (op == prev_op) ++e;
        else {
            if (l >= cap) return -1;
            out[cap - 1 - l] = ((uint32_t)e << 4) | (uint32_t)prev_op; ++l;
            prev_op = op; e = 1;
        }
    }
    if (op == 0) {
        if (l >= cap) return -1;
        out[cap - 1 - l] = ((uint32_t)(e + 1) << 4); ++l;
    } else {
        if (l + 1 >= cap) return -1;
        out[cap - 1 - l] = ((uint32_t)e << 4) | (uint32_t)op; ++l;
        out[cap - 1 - l] = (1u << 4); ++l;
    }
    // entries were produced end-to-start; they sit reversed at the tail, i.e. already in forward order
    for (int k = 0; k < l; ++k) out[k] = out[cap - l + k];
    return l;
}

__global__ void __launch_bounds__(PAIRS_THREADS) sw_pairs_kernel(PairsParams p) {
    __shared__ SwLut lut;
    sw_build_lut(&lut, c_mat25, threadIdx.x, blockDim.x);
    __syncthreads();
    const int nthreads = gridDim.x * blockDim.x;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t *bnd = p.bnd + tid;
    for (int64_t k = tid; k < p.npairs; k += nthreads) {
        const int qi = p.qidx[k], ti = p.tidx[k];
        const int8_t *q = p.qbuf + p.qoff[qi];
        const int8_t *t = p.tbuf + p.toff[ti];
        const int m = (int)(p.qoff[qi + 1] - p.qoff[qi]);
        const int n = (int)(p.toff[ti + 1] - p.toff[ti]);
        int32_t *o = p.out + k * 8;
        auto clampq = [](int c) { return (c < 0 || c > 4) ? 4 : c; };
        auto rc_f = [&](int j) { return j < m ? clampq(q[j]) : SW_CODE_GHOST; };
        auto cc_f = [&](int i) { return clampq(t[i]); };
        int dummy_c, dummy_r;
        int score, score2 = 0, ref_end2 = -1;
        bool word = false;
        if (p.flags & TREDSW_SCORE2) {
            // per-column maxima go to a thread-private contiguous array; the striped kernels' ghost
            // padding rows (16 lanes in the 8-bit kernel, 8 in the 16-bit one) are part of them
            int32_t *cmv = p.colmax + (int64_t)tid * p.max_n;
            int mpad = ((m + 15) / 16) * 16;
            if (p.flags & TREDSW_FORCE_WORD) mpad = ((m + 7) / 8) * 8;
            score = sw_sweep<PAIRS_W, 0, true>(m, mpad, n, rc_f, cc_f, &lut, bnd, nthreads, p.go, p.ge, 0,
                                               &dummy_c, &dummy_r, cmv);
            word = (score + p.bias >= 255) || (p.flags & TREDSW_FORCE_WORD);
            if (word && !(p.flags & TREDSW_FORCE_WORD)) {
                mpad = ((m + 7) / 8) * 8;
                score = sw_sweep<PAIRS_W, 0, true>(m, mpad, n, rc_f, cc_f, &lut, bnd, nthreads, p.go, p.ge,
                                                   0, &dummy_c, &dummy_r, cmv);
            }
        } else {
            score = sw_sweep<PAIRS_W, 0, false>(m, m, n, rc_f, cc_f, &lut, bnd, nthreads, p.go, p.ge, 0,
                                                &dummy_c, &dummy_r, nullptr);
            word = (score + p.bias >= 255) || (p.flags & TREDSW_FORCE_WORD);
        }
        if (score <= 0) {
            // nothing aligned (src/ssw.c byte path: end_ref = -1, end_read = 0)
            o[0] = 0; o[1] = -1; o[2] = -1; o[3] = 0; o[4] = 0; o[5] = 0; o[6] = (p.flags & TREDSW_SCORE2) ? 0 : -1; o[7] = 0;
            continue;
        }
        int end_ref, end_read;
        sw_sweep<PAIRS_W, 1, false>(m, m, n, rc_f, cc_f, &lut, bnd, nthreads, p.go, p.ge, score, &end_ref,
                                    &end_read, nullptr);
        if (p.flags & TREDSW_SCORE2) {
            // src/ssw.c:315-340 (byte) / 519-542 (word); caller's maskLen convention src/ssw_wrap.py:198-201
            const int32_t *cmv = p.colmax + (int64_t)tid * p.max_n;
            const int mask = m > 30 ? m / 2 : 15;
            int s2 = 0, r2 = 0;
            int edge = (end_ref - mask) > 0 ? (end_ref - mask) : 0;
            for (int i = 0; i < edge; ++i) if (cmv[i] > s2) { s2 = cmv[i]; r2 = i; }
            edge = (end_ref + mask) > n ? n : (end_ref + mask);
            for (int i = edge + (word ? 0 : 1); i < n; ++i) if (cmv[i] > s2) { s2 = cmv[i]; r2 = i; }
            score2 = s2; ref_end2 = r2;
        }
        int ref_begin = -1, query_begin = -1;
        if (!(p.flags & TREDSW_NO_BEGIN)) {
            auto rc_r = [&](int j) { return clampq(q[end_read - j]); };
            auto cc_r = [&](int i) { return clampq(t[end_ref - i]); };
            int ci, rj;
            sw_sweep<PAIRS_W, 1, false>(end_read + 1, end_read + 1, end_ref + 1, rc_r, cc_r, &lut, bnd, nthreads,
                                        p.go, p.ge, score, &ci, &rj, nullptr);
            ref_begin = end_ref - ci;
            query_begin = end_read - rj;
        }
        int cigar_len = 0;
        if ((p.flags & TREDSW_CIGAR) && ref_begin >= 0) {
            cigar_len = banded_cigar(t + ref_begin, q + query_begin, end_ref - ref_begin + 1,
                                     end_read - query_begin + 1, score, p.go, p.ge,
                                     p.hbuf + (int64_t)tid * p.hbuf_per_thread, p.hbuf_per_thread,
                                     p.dirs + (int64_t)tid * p.dirs_per_thread, p.dirs_per_thread,
                                     p.cigar_out + k * p.cigar_cap, p.cigar_cap);
        }
        o[0] = score; o[1] = ref_begin; o[2] = end_ref; o[3] = query_begin; o[4] = end_read;
        o[5] = score2; o[6] = ref_end2; o[7] = cigar_len;
    }
}

}  // namespace

extern "C" int tredsw_align_pairs(tredsw_ctx *ctx, const int8_t *qbuf, const int64_t *qoff, int32_t nq,
                                  const int8_t *tbuf, const int64_t *toff, int32_t nt,
                                  const int32_t *qidx, const int32_t *tidx, int64_t npairs,
                                  const int8_t *mat25, int gap_open, int gap_extend, uint32_t flags,
                                  int32_t *out, uint32_t *cigar_out, int32_t cigar_cap) {
    if (!ctx) { tredsw_set_error("null context"); return TREDSW_ERR_ARG; }
    if (npairs < 0 || nq < 0 || nt < 0 || !mat25) { tredsw_set_error("bad arguments"); return TREDSW_ERR_ARG; }
    if (npairs == 0) return TREDSW_OK;
    if (dev_ptrs(flags)) {
        tredsw_set_error("tredsw_align_pairs: device-pointer mode needs host-visible offsets; use host pointers");
        return TREDSW_ERR_UNSUPPORTED;
    }
    if ((flags & TREDSW_CIGAR) && (!cigar_out || cigar_cap < 4)) { tredsw_set_error("cigar buffer missing"); return TREDSW_ERR_ARG; }
    std::lock_guard<std::mutex> lock(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    // sizes
    int max_m = 0, max_n = 0;
    for (int i = 0; i < nq; ++i) { int l = (int)(qoff[i + 1] - qoff[i]); if (l > max_m) max_m = l; }
    for (int i = 0; i < nt; ++i) { int l = (int)(toff[i + 1] - toff[i]); if (l > max_n) max_n = l; }
    if (max_m > 65000 || max_n > 1000000) { tredsw_set_error("sequence too long"); return TREDSW_ERR_UNSUPPORTED; }
    const int max_rows = ((max_m + 15) / 16) * 16 + 16;
    int bias = 0;
    for (int i = 0; i < 25; ++i) if (mat25[i] < bias) bias = mat25[i];
    bias = -bias;

    int blocks = (int)((npairs + PAIRS_THREADS - 1) / PAIRS_THREADS);
    int max_blocks = ctx->sm_count * 8;
    if (flags & TREDSW_CIGAR) {
        // traceback scratch is large (direction bytes for the widest band): cap it at ~3 GiB
        int64_t wmax = 4 * (int64_t)(max_m > max_n ? max_m : max_n) + 16;
        int64_t per_thread = wmax * max_m * 3 + 16;
        int64_t fit = ((int64_t)3 << 30) / (per_thread * PAIRS_THREADS);
        if (fit < 1) { tredsw_set_error("sequences too long for CIGAR scratch"); return TREDSW_ERR_UNSUPPORTED; }
        if (max_blocks > fit) max_blocks = (int)fit;
    }
    if (blocks > max_blocks) blocks = max_blocks;
    const int nthreads = blocks * PAIRS_THREADS;

    PairsParams p{};
    int rc;
    if ((rc = stage_in(ctx, ctx->d_q, qbuf, (size_t)qoff[nq], flags, &p.qbuf))) return rc;
    if ((rc = stage_in(ctx, ctx->d_qoff, qoff, (size_t)nq + 1, flags, &p.qoff))) return rc;
    if ((rc = stage_in(ctx, ctx->d_t, tbuf, (size_t)toff[nt], flags, &p.tbuf))) return rc;
    if ((rc = stage_in(ctx, ctx->d_toff, toff, (size_t)nt + 1, flags, &p.toff))) return rc;
    if ((rc = stage_in(ctx, ctx->d_qidx, qidx, (size_t)npairs, flags, &p.qidx))) return rc;
    if ((rc = stage_in(ctx, ctx->d_tidx, tidx, (size_t)npairs, flags, &p.tidx))) return rc;
    if ((rc = ctx->d_out.ensure((size_t)npairs * 8 * sizeof(int32_t)))) return rc;
    if ((rc = ctx->d_scratch.ensure((size_t)max_rows * nthreads * sizeof(uint32_t)))) return rc;
    p.out = ctx->d_out.as<int32_t>();
    p.bnd = ctx->d_scratch.as<uint32_t>();
    p.max_rows = max_rows; p.max_n = max_n;
    if (flags & TREDSW_SCORE2) {
        if ((rc = ctx->d_misc.ensure((size_t)max_n * nthreads * sizeof(int32_t)))) return rc;
        p.colmax = ctx->d_misc.as<int32_t>();
    }
    if (flags & TREDSW_CIGAR) {
        // banded_sw scratch: the band doubles until the score is reproduced, at worst until it covers
        // the whole sub-matrix (< 2 * max dimension)
        int64_t wmax = 4 * (int64_t)(max_m > max_n ? max_m : max_n) + 16;
        p.hbuf_per_thread = (int)(3 * (wmax + 1));
        p.dirs_per_thread = wmax * max_m * 3 + 16;
        if ((rc = ctx->d_work.ensure((size_t)p.hbuf_per_thread * nthreads * sizeof(int32_t)))) return rc;
        if ((rc = ctx->d_cigar.ensure((size_t)npairs * cigar_cap * sizeof(uint32_t) +
                                      (size_t)p.dirs_per_thread * nthreads))) return rc;
        p.hbuf = ctx->d_work.as<int32_t>();
        p.cigar_out = ctx->d_cigar.as<uint32_t>();
        p.dirs = reinterpret_cast<int8_t *>(p.cigar_out + (size_t)npairs * cigar_cap);
        p.cigar_cap = cigar_cap;
    }
    p.npairs = npairs; p.go = gap_open; p.ge = gap_extend; p.bias = bias; p.flags = flags;
    CUDA_TRY(cudaMemcpyToSymbolAsync(c_mat25, mat25, 25, 0, cudaMemcpyHostToDevice, ctx->stream));
    sw_pairs_kernel<<<blocks, PAIRS_THREADS, 0, ctx->stream>>>(p);
    CUDA_TRY(cudaGetLastError());
    ctx->launches += 1;
    CUDA_TRY(cudaMemcpyAsync(out, p.out, (size_t)npairs * 8 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (flags & TREDSW_CIGAR)
        CUDA_TRY(cudaMemcpyAsync(cigar_out, p.cigar_out, (size_t)npairs * cigar_cap * sizeof(uint32_t),
                                 cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return TREDSW_OK;
}

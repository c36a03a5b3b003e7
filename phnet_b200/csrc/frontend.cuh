// frontend.cuh -- the callers either side of the op, for a whole clip at once:
//
//   phnms_prepare_kernel   get_lanes' preparation, libs/models/Router4OL.py:447-458 / RouterV4.py:404-418:
//                          score = softmax(logits)[1]; keep score >= conf_threshold; drop the theta (and, VIL, the
//                          invalid-length) column; scale start_x and the x offsets to pixels, the length to strips.
//                          Compaction keeps the prior order, so the indices nms returns mean the same thing.
//   (lane NMS)             phnms_forward_f32 on the compacted rows with n_valid
//   phnms_gather_kernel    predictions[keep] and the rounding of the length column(s), Router4OL.py:465-470
//
// The reference does this once per frame in Python with ~10 small kernels and two host syncs (the boolean-mask
// compaction and `keep[:num_to_keep]`); here a clip of T frames takes 4 launches and no sync.
//
// softmax is the two-element case of ATen's persistent warp softmax (softmax_warp_forward): max = (e1 < e0) ? e0 : e1,
// x_i = expf(e_i - max), out_1 = x_1 / (x_1 + x_0) -- checked bit for bit against torch.softmax on the GPU (tests).
#pragma once
#include "common.cuh"

namespace phnms {

constexpr int kPrepThreads = 256;

// pred [T, A, hdr + n_off] (hdr = 6: OpenLane-V rows, hdr = 7: VIL-100 rows with the invalid-length column)
__global__ void __launch_bounds__(kPrepThreads) phnms_prepare_kernel(const float *__restrict__ pred, int A, int hdr,
                                                                    int n_off, float conf_thr, float img_w_m1,
                                                                    float n_strips, float *__restrict__ cprops,
                                                                    float *__restrict__ cscores, int *__restrict__ src,
                                                                    int *__restrict__ n_valid,
                                                                    unsigned char *__restrict__ keep_inds) {
    extern __shared__ int prep_src[];          // [A] prior index of compacted row r
    __shared__ int warp_cnt[kPrepThreads / 32];
    __shared__ int base_s;
    const long long t = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = hdr + n_off, P = 5 + n_off;
    const float *frame = pred + (size_t)t * A * C;
    if (tid == 0) base_s = 0;
    __syncthreads();
    for (int a0 = 0; a0 < A; a0 += kPrepThreads) {
        const int a = a0 + tid;
        bool keepf = false;
        float score = 0.0f;
        if (a < A) {
            const float e0 = frame[(size_t)a * C], e1 = frame[(size_t)a * C + 1];
            const float mx = (e1 < e0) ? e0 : e1;
            const float x0 = expf(__fsub_rn(e0, mx)), x1 = expf(__fsub_rn(e1, mx));
            score = __fdiv_rn(x1, __fadd_rn(x1, x0));
            keepf = score >= conf_thr;
            keep_inds[(size_t)t * A + a] = keepf ? 1 : 0;
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, keepf);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int off = base_s;
        for (int w = 0; w < warp; ++w) off += warp_cnt[w];
        if (keepf) {
            const int pos = off + __popc(bal & ((1u << lane) - 1u));
            prep_src[pos] = a;
            cscores[(size_t)t * A + pos] = score;
            src[(size_t)t * A + pos] = a;
        }
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < kPrepThreads / 32; ++w) tot += warp_cnt[w];
            base_s += tot;
        }
        __syncthreads();
    }
    const int nv = base_s;
    if (tid == 0) n_valid[t] = nv;
    // one warp per kept row: coalesced copy with the column drop and the scalings
    for (int r = warp; r < nv; r += kPrepThreads / 32) {
        const float *row = frame + (size_t)prep_src[r] * C;
        float *dst = cprops + ((size_t)t * A + r) * P;
        for (int i = lane; i < P; i += 32) {
            float v;
            if (i < 3) v = row[i];
            else if (i == 3) v = __fmul_rn(row[3], img_w_m1);
            else if (i == 4) v = __fmul_rn(row[5], n_strips);
            else v = __fmul_rn(row[hdr + (i - 5)], img_w_m1);
            dst[i] = v;
        }
    }
}

// out_rows [T, K, hdr + n_off]: predictions[keep] with columns 5 .. hdr-1 replaced by round(col * n_strips);
// out_index [T, K]: the kept priors' indices in the ORIGINAL (unfiltered) frame; both zero past num[t].
__global__ void __launch_bounds__(128) phnms_gather_kernel(const float *__restrict__ pred, int A, int hdr, int n_off,
                                                          float n_strips, const long long *__restrict__ keep,
                                                          const long long *__restrict__ num,
                                                          const int *__restrict__ src, int K,
                                                          float *__restrict__ out_rows,
                                                          long long *__restrict__ out_index) {
    const long long t = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int C = hdr + n_off;
    const long long nk = num[t];
    for (int k = warp; k < K; k += 4) {
        float *dst = out_rows + ((size_t)t * K + k) * C;
        if (k < nk) {
            const int a = src[(size_t)t * A + keep[(size_t)t * A + k]];
            const float *row = pred + ((size_t)t * A + a) * C;
            for (int i = lane; i < C; i += 32) dst[i] = (i >= 5 && i < hdr) ? rintf(__fmul_rn(row[i], n_strips)) : row[i];
            if (lane == 0) out_index[(size_t)t * K + k] = a;
        } else {
            for (int i = lane; i < C; i += 32) dst[i] = 0.0f;
            if (lane == 0) out_index[(size_t)t * K + k] = 0;
        }
    }
}

}  // namespace phnms

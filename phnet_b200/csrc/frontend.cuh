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
#include "select.cuh"

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

// ---- get_lanes in ONE launch -----------------------------------------------------------------------
// get_lanes uses only `keep[:num_to_keep]` of the op's three results (`keep, num_to_keep, _ = nms(...)`, Router4OL.py:460-465):
// parent_object_index -- the part of the op that touches every proposal -- is never looked at.  What get_lanes needs is exactly
// what the select step computes: the greedy scan over the proposals in rank order until top_k lanes are kept.  So for a clip the
// whole of get_lanes' tensor work is one kernel, one warp per frame, reading the raw head output `pred` directly:
//   * score = softmax(logits)[1] and the confidence filter while the rank keys are built (the filtered priors simply get no
//     key; compaction keeps the prior order, so ordering by (key, prior index) IS the order of the compacted frame -- and when
//     <= 32 priors survive, ATen's unstable bitonic network is replayed on the compacted scores, as torch's sort would);
//   * the column drop and the pixel / strip scalings are applied to the few candidate rows when they are fetched -- the
//     compacted, rescaled copy of the frame that the unfused pipeline writes to HBM and reads back never exists;
//   * draws continue until top_k lanes are kept or the frame is exhausted (no cap: at most A priors), so there is no resume pass;
//   * predictions[keep] with the length column(s) rounded is written straight from `pred`.
// Same arithmetic as phnms_prepare_kernel / phnms_select_kernel / phnms_gather_kernel, bit for bit (tests/test_get_lanes_gpu.py).
struct GetLanesFusedParams {
    const float *pred;
    long long T;
    int A, hdr, n_off, top_k, sort_model;
    float conf_thr, img_w_m1, n_strips, thr;
    float *out_rows;
    long long *out_num;
    long long *out_index;
    unsigned char *keep_inds;
};

__host__ __device__ inline int get_lanes_warp_words(int A, int n_off, int top_k) {
    const int P4 = (5 + n_off + 3) & ~3, G = (A + 31) / 32;
    const int w = top_k * (8 + P4) + kSelBatch * (P4 | 1) + 2 * kSelBatch + kSelBatch + 32 * (G | 1) + G + 8 + 64;
    return (w + 3) & ~3;
}

__global__ void __launch_bounds__(kSelWarps * 32) phnms_get_lanes_fused_kernel(const GetLanesFusedParams gp) {
    extern __shared__ __align__(16) unsigned char smem_gl[];
    __shared__ float bit_key[kSelWarps][32];
    __shared__ int bit_val[kSelWarps][32];
    __shared__ int bit_ok[kSelWarps][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const long long t = (long long)blockIdx.x * nw + warp;
    if (t >= gp.T) return;
    const int A = gp.A, n_off = gp.n_off, hdr = gp.hdr, C = hdr + n_off, P = 5 + n_off, P4 = (P + 3) & ~3, slot_words = 8 + P4, bp = P4 | 1;
    const int top_k = gp.top_k, G = (A + 31) / 32, pitch = G | 1;
    uint32_t *slots = reinterpret_cast<uint32_t *>(smem_gl) + (size_t)warp * get_lanes_warp_words(A, n_off, top_k);
    float *brow = reinterpret_cast<float *>(slots + top_k * slot_words);
    int *bse = reinterpret_cast<int *>(brow + kSelBatch * bp);
    uint32_t *adj = reinterpret_cast<uint32_t *>(bse + 2 * kSelBatch);
    uint32_t *kb = adj + kSelBatch;                    // [32][pitch] rank keys
    uint32_t *vm = kb + 32 * pitch;                    // [G] survivors of priors 32 q .. 32 q + 31
    int *kept_idx = reinterpret_cast<int *>(vm + G);   // [8] prior index of the k-th kept lane
    float *cs = reinterpret_cast<float *>(kept_idx + 8);   // [32] scores of the first 32 survivors, in prior order
    int *ca = reinterpret_cast<int *>(cs + 32);            // [32] their prior indices
    const float *frame = gp.pred + (size_t)t * A * C;
    const bool nan_first = gp.sort_model == 1;

    // ---- scores, the confidence filter, rank keys -----------------------------------------------------------------------------
    u64 gmin = kNone64;
    int n = 0;
    for (int q = 0; q < G; ++q) {
        const int a = lane + 32 * q;
        bool keepf = false;
        float score = 0.0f;
        if (a < A) {
            const float e0 = frame[(size_t)a * C], e1 = frame[(size_t)a * C + 1];
            const float mx = (e1 < e0) ? e0 : e1;
            const float x0 = expf(__fsub_rn(e0, mx)), x1 = expf(__fsub_rn(e1, mx));
            score = __fdiv_rn(x1, __fadd_rn(x1, x0));
            keepf = score >= gp.conf_thr;
            gp.keep_inds[(size_t)t * A + a] = keepf ? 1 : 0;
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, keepf);
        uint32_t k = 0xffffffffu;
        if (keepf) {
            k = key_desc(score, nan_first);
            gmin = min(gmin, ((u64)k << 32) | (uint32_t)a);
            const int pos = n + __popc(bal & ((1u << lane) - 1u));     // position in the compacted frame
            if (pos < 32) {
                cs[pos] = score;
                ca[pos] = a;
            }
        }
        kb[lane * pitch + q] = k;
        if (lane == 0) vm[q] = bal;
        n += __popc(bal);
    }
    __syncwarp();
    const bool bitonic = gp.sort_model == 0 && n <= 32 && n >= 2;
    u64 sorted = kNone64;
    if (bitonic) {          // ATen bitonicSortKVInPlace<block_dim_x = 16> on the compacted scores (see topm.cuh)
        float *bk = bit_key[warp];
        int *bv = bit_val[warp], *bo = bit_ok[warp];
        bo[lane] = lane < n;
        bk[lane] = lane < n ? cs[lane] : 0.0f;
        bv[lane] = lane < n ? lane : 0;
        __syncwarp();
        for (unsigned size = 2; size <= 32; size *= 2) {
            const bool flag = (size != 32) && ((lane & (size / 2)) != 0);
            for (unsigned stride = size / 2; stride > 0; stride /= 2) {
                if (lane < 16) {
                    const unsigned pa = 2 * lane - (lane & (stride - 1)), pb = pa + stride;
                    const float ka = bk[pa], kbv = bk[pb];
                    const int oa = bo[pa], ob = bo[pb];
                    const bool sw = (gt_nan(ka, kbv) && oa) || !ob;
                    if (sw == flag) {
                        const int va = bv[pa], vb = bv[pb];
                        bk[pa] = kbv; bk[pb] = ka;
                        bv[pa] = vb; bv[pb] = va;
                        bo[pa] = ob; bo[pb] = oa;
                    }
                }
                __syncwarp();
            }
        }
        if (lane < n) sorted = ((u64)(uint32_t)lane << 32) | (uint32_t)ca[bv[lane]];   // rank = sorted position, index = the prior
    }

    // ---- the greedy scan over the proposals in rank order, batch by batch ----------------------------------------------------------
    int nk = 0, drawn = 0;
    while (n > 0) {
        u64 myc = kNone64;
        int nb = 0;
        if (bitonic) {
            nb = min(kSelBatch, n - drawn);
            const u64 v = __shfl_sync(0xffffffffu, sorted, (drawn + lane) & 31);
            if (lane < nb) myc = v;
        } else {
            nb = select_draw_batch(gmin, kb, pitch, G, lane, myc, [&](int i, int q) { return ((vm[q] >> (i & 31)) & 1u) != 0u; });
        }
        if (nb == 0) break;
        {   // candidate rows straight from `pred`: theta (and the invalid-length column) dropped, start_x / x in pixels, length in strips
            float rv[kSelBatch][3];
#pragma unroll
            for (int j = 0; j < kSelBatch; ++j) {
                const u64 kj = __shfl_sync(0xffffffffu, myc, j);
                const float *row = frame + (size_t)(j < nb ? (uint32_t)kj : 0u) * C;
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    const int i = lane + 32 * u;
                    float v = 0.0f;
                    if (j < nb && i < P) {
                        if (i < 3) v = row[i];
                        else if (i == 3) v = __fmul_rn(row[3], gp.img_w_m1);
                        else if (i == 4) v = __fmul_rn(row[5], gp.n_strips);
                        else v = __fmul_rn(row[hdr + (i - 5)], gp.img_w_m1);
                    }
                    rv[j][u] = v;
                }
            }
#pragma unroll
            for (int j = 0; j < kSelBatch; ++j)
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    const int i = lane + 32 * u;
                    if (j < nb && i < P4) brow[j * bp + i] = rv[j][u];
                }
        }
        __syncwarp();
        nk = select_scan_batch(nb, myc, nk, top_k, n_off, gp.thr, slots, brow, bse, adj, lane, [&](int k, u64 Ka) {
            if (lane == 0) kept_idx[k] = (int)(uint32_t)Ka;
        });
        drawn += nb;
        if (nk == top_k || drawn >= n) break;
    }
    __syncwarp();

    // ---- predictions[keep], the length column(s) rounded (Router4OL.py:465-470) ------------------------------------------------------
    for (int k = 0; k < top_k; ++k) {
        float *dst = gp.out_rows + ((size_t)t * top_k + k) * C;
        if (k < nk) {
            const int a = kept_idx[k];
            const float *row = frame + (size_t)a * C;
            for (int i = lane; i < C; i += 32) dst[i] = (i >= 5 && i < hdr) ? rintf(__fmul_rn(row[i], gp.n_strips)) : row[i];
            if (lane == 0) gp.out_index[(size_t)t * top_k + k] = a;
        } else {
            for (int i = lane; i < C; i += 32) dst[i] = 0.0f;
            if (lane == 0) gp.out_index[(size_t)t * top_k + k] = 0;
        }
    }
    if (lane == 0) gp.out_num[t] = (long long)nk;
}

// ---- the training-side line IoU -----------------------------------------------------------
// libs/utils/dynamic_assign.py:5-36 `line_iou(pred, target, img_w, length, aligned)`: every x offset is widened to a
// segment of radius `length`; IoU = sum of per-offset overlaps / (sum of per-offset unions + 1e-9), offsets whose target is
// outside [0, img_w) contributing nothing.  aligned: pair i of two equally long lists; otherwise the full
// [num_pred, num_target] matrix the dynamic-k assignment consumes (:83-125).  fp32 throughout like the reference; the
// sums run in ascending offset order (torch's vectorised reduction order differs: parity is to 1e-5 relative, stated in
// the test).  torch.min / torch.max propagate NaN, fminf / fmaxf do not: emulated.
__device__ __forceinline__ float tmin(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : fminf(a, b); }
__device__ __forceinline__ float tmax(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : fmaxf(a, b); }

__global__ void __launch_bounds__(128) phnms_line_iou_kernel(const float *__restrict__ pred, const float *__restrict__ target,
                                                            int num_pred, int num_target, int n_off, float img_w,
                                                            float length, int aligned, float *__restrict__ out) {
    extern __shared__ float liou_smem[];                 // [128][n_off + 1] pred tile, then [tile_t][n_off] targets
    const int pitch = n_off + 1;
    float *ps = liou_smem;
    float *ts = liou_smem + 128 * pitch;
    const int p0 = blockIdx.x * 128, tid = threadIdx.x;
    const int np = min(128, num_pred - p0);
    for (int i = tid; i < np * n_off; i += 128) ps[(i / n_off) * pitch + (i % n_off)] = pred[(size_t)p0 * n_off + i];
    if (aligned) {
        __syncthreads();
        if (tid < np) {
            const float *t = target + (size_t)(p0 + tid) * n_off;
            float so = 0.0f, su = 0.0f;
            for (int i = 0; i < n_off; ++i) {
                const float pv = ps[tid * pitch + i], tv = t[i];
                const float px1 = __fsub_rn(pv, length), px2 = __fadd_rn(pv, length);
                const float tx1 = __fsub_rn(tv, length), tx2 = __fadd_rn(tv, length);
                const bool invalid = (tv < 0.0f) || (tv >= img_w);
                const float o = __fsub_rn(tmin(px2, tx2), tmax(px1, tx1)), u = __fsub_rn(tmax(px2, tx2), tmin(px1, tx1));
                so = __fadd_rn(so, invalid ? 0.0f : o);
                su = __fadd_rn(su, invalid ? 0.0f : u);
            }
            out[p0 + tid] = __fdiv_rn(so, __fadd_rn(su, 1e-9f));
        }
        return;
    }
    for (int t0 = 0; t0 < num_target; t0 += 32) {          // targets in tiles of 32 (a frame has a handful of lanes)
        const int nt = min(32, num_target - t0);
        __syncthreads();
        for (int i = tid; i < nt * n_off; i += 128) ts[i] = target[(size_t)t0 * n_off + i];
        __syncthreads();
        if (tid < np) {
            for (int t = 0; t < nt; ++t) {
                float so = 0.0f, su = 0.0f;
                for (int i = 0; i < n_off; ++i) {
                    const float pv = ps[tid * pitch + i], tv = ts[t * n_off + i];
                    const float px1 = __fsub_rn(pv, length), px2 = __fadd_rn(pv, length);
                    const float tx1 = __fsub_rn(tv, length), tx2 = __fadd_rn(tv, length);
                    const bool invalid = (tv < 0.0f) || (tv >= img_w);
                    const float o = __fsub_rn(tmin(px2, tx2), tmax(px1, tx1)), u = __fsub_rn(tmax(px2, tx2), tmin(px1, tx1));
                    so = __fadd_rn(so, invalid ? 0.0f : o);
                    su = __fadd_rn(su, invalid ? 0.0f : u);
                }
                out[(size_t)(p0 + tid) * num_target + t0 + t] = __fdiv_rn(so, __fadd_rn(su, 1e-9f));
            }
        }
    }
}

// ---- predictions_to_pred on the device ----------------------------------------------------
// The tensor part of libs/models/Router4OLV2.py:363-404 (hdr == 6) and RouterV4.py:349-392 (hdr == 7) for every kept lane
// of a clip: start / end rounding, the "extend to the bottom" mask, the -2 fill, selection of the points with x >= 0, the
// flip, the y rescale -- everything up to the `Lane(points=...)` constructor, which (a scipy spline) stays on the host.
// One warp per (frame, slot).  Python semantics reproduced: round() is round-half-even on the double of the fp32 value
// (rint), slices with negative bounds wrap like Python's (`lane_xs[end + 1:]` with end + 1 < 0), and the mask is computed
// from the row BEFORE the -2 fills.
//   rows     [T, K, hdr + n_off] fp32: get_lanes output (length column(s) already rounded, Router4OL.py:470)
//   num      [T] int64 kept lanes per frame
//   prior_ys [n_off] fp64: `torch.linspace(1, 0, n_off)` as the model holds it (Router4OLV2.py:61), widened to double
//   points   [T, K, n_off, 2] fp64: (x, y) of the lane's points in the order `Lane.points` has them (flipped), zero padded
//   npoints  [T, K] int32: points of the lane; 0 where the reference `continue`s (<= 1 point) or the slot is empty
//   meta     [T, K, 3] fp32: start_x, start_y, conf  (lane[3], lane[2], lane[1], the Lane metadata)
__device__ __forceinline__ int py_slice_index(int i, int n) { return i < 0 ? max(n + i, 0) : min(i, n); }

__global__ void __launch_bounds__(128) phnms_decode_kernel(const float *__restrict__ rows, const long long *__restrict__ num,
                                                          int K, int hdr, int n_off, const double *__restrict__ prior_ys,
                                                          double ori_img_h, double cut_height, double *__restrict__ points,
                                                          int *__restrict__ npoints, float *__restrict__ meta,
                                                          long long total) {
    __shared__ float xs_s[4][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long slot = (long long)blockIdx.x * 4 + warp;
    if (slot >= total) return;
    const long long t = slot / K;
    const int k = (int)(slot - t * K);
    const int C = hdr + n_off, n_strips = n_off - 1;
    const float *row = rows + (size_t)slot * C;
    double *out = points + (size_t)slot * n_off * 2;
    float *xs = xs_s[warp];
    for (int i = lane; i < n_off; i += 32) {
        xs[i] = row[hdr + i];
        out[2 * i] = 0.0;
        out[2 * i + 1] = 0.0;
    }
    __syncwarp();
    if (k >= num[t]) {
        if (lane == 0) npoints[slot] = 0;
        if (lane < 3) meta[slot * 3 + lane] = 0.0f;
        return;
    }
    // start = min(max(0, int(round(lane[2] * n_strips))), n_strips) [+ invalid_len]; end = min(start + length - 1, n_off - 1)
    int start = (int)rint((double)row[2] * (double)n_strips);
    start = min(max(0, start), n_strips);
    if (hdr == 7) start += (int)rint((double)row[6]);
    const int length = (int)rint((double)row[5]);
    const int end = min(start + length - 1, n_off - 1);
    const int fill_from = py_slice_index(end + 1, n_off);   // lane_xs[end + 1:] = -2
    const int head_len = py_slice_index(start, n_off);       // lane_xs[:start]
    int cnt_after[3];                                         // valid points with a larger index, per 32-chunk
    bool valid[3];
    uint32_t bal[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int i = lane + 32 * c;
        bool v = false;
        if (i < n_off) {
            bool masked = i >= fill_from;
            if (i < head_len) {
                if (hdr == 7) {
                    masked = true;                            // RouterV4.py:372 lane_xs[:start] = -2
                } else {                                      // Router4OLV2.py:382-385: keep only the run of in-image x that reaches `start`
                    bool run = true;
                    for (int j = i; j < head_len; ++j) run = run && (xs[j] >= 0.0f) && (xs[j] <= 1.0f);
                    masked = masked || !run;
                }
            }
            v = !masked && xs[i] >= 0.0f;                     // lane_xs >= 0 (after the -2 fills)
        }
        valid[c] = v;
        bal[c] = __ballot_sync(0xffffffffu, v);
    }
    const int tot = __popc(bal[0]) + __popc(bal[1]) + __popc(bal[2]);
    cnt_after[2] = 0;
    cnt_after[1] = __popc(bal[2]);
    cnt_after[0] = __popc(bal[2]) + __popc(bal[1]);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (valid[c]) {
            const int i = lane + 32 * c;
            const int pos = cnt_after[c] + __popc(bal[c] & ~((2u << lane) - 1u));   // flipped: larger indices come first
            double y = prior_ys[i];
            if (hdr == 7) y = (y * (ori_img_h - cut_height) + cut_height) / ori_img_h;  // RouterV4.py:378
            out[2 * pos] = (double)xs[i];
            out[2 * pos + 1] = y;
        }
    }
    if (lane == 0) npoints[slot] = tot <= 1 ? 0 : tot;       // `if len(lane_xs) <= 1: continue`
    if (lane < 3) meta[slot * 3 + lane] = row[3 - lane];     // start_x = lane[3], start_y = lane[2], conf = lane[1]
}

}  // namespace phnms

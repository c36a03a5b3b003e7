// assign.cuh -- the dynamic-k assignment of PHNet's training code on the device (second half).
//
// Reference: libs/utils/dynamic_assign.py:83-125 `dynamic_k_assign(cost, pair_wise_ious)` (the same body lives in
// libs/utils/dynamic_assignV2.py:372-405 with max_topk / min_topk, and :327-370 `dynamic_k_assign_CF` with binarised IoUs,
// one candidate and a minimum of 0).  In the reference it is ~10 + 3 per ground truth small torch launches and two host syncs
// (`dynamic_ks[gt_idx].item()`, `nonzero`) per image; here one CTA per image does all of it:
//
//   1. dynamic k per ground truth (:97-101): the n_candidate_k largest IoUs of the column (negatives clipped to 0, or
//      binarised at a threshold), summed in descending order in fp32, truncated to int, clamped from below;
//   2. column by column (:105-111): the k smallest entries of cost4match[:, gt] -- one CTA-wide arg-min per pick, rows taken
//      by an earlier column count as INFINITY (987654.0, :3) -- are matched and taken;
//   3. a prior matched to several ground truths keeps the one of least cost (:116-120);
//   4. prior_idx = matched priors in ascending order, gt_idx = their ground truth (:122-124).
//
// One prior per thread (num_priors <= 1024; PHNet has 240).  Ties: torch.topk leaves the order among equal values open; this
// kernel takes the lowest index.  The fixtures use tie-free costs (and the one structural tie -- several taken rows, all
// INFINITY, picked again as a complete set -- which is unambiguous).  NaN inputs are not supported.
#pragma once
#include "common.cuh"

namespace phnms {

constexpr int kAssignMaxPriors = 1024;
constexpr int kAssignMaxGt = 1024;
constexpr int kAssignMaxCand = 8;
constexpr float kAssignInfinity = 987654.0f;   // dynamic_assign.py:3

// ascending-u32 order == ascending float order (no NaN)
__device__ __forceinline__ uint32_t float_order_key(float v) {
    const uint32_t u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(kAssignMaxPriors) phnms_dynamic_k_assign_kernel(
    const float *__restrict__ cost_all, const float *__restrict__ iou_all, int num_priors, int num_gt, int n_cand, int min_k,
    int binarize, float binarize_at, long long *__restrict__ prior_idx, long long *__restrict__ gt_idx,
    long long *__restrict__ count) {
    extern __shared__ int assign_smem[];          // [num_gt] dynamic ks
    __shared__ u64 wred[2][32];
    __shared__ int wcount[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const size_t b = blockIdx.x;
    const float *cost = cost_all + b * (size_t)num_priors * num_gt;
    const float *iou = iou_all + b * (size_t)num_priors * num_gt;
    int *ks = assign_smem;

    // ---- 1. dynamic k per ground truth: one warp per column ------------------------------------------------------------------
    for (int g = warp; g < num_gt; g += nwarps) {
        float top[kAssignMaxCand];                // this lane's largest values of the column, descending; -1 = empty
#pragma unroll
        for (int r = 0; r < kAssignMaxCand; ++r) top[r] = -1.0f;
        for (int p = lane; p < num_priors; p += 32) {
            float v = iou[(size_t)p * num_gt + g];
            v = binarize ? (v >= binarize_at ? 1.0f : 0.0f) : (v < 0.0f ? 0.0f : v);   // :95 / dynamic_k_assign_CF :340-341
#pragma unroll
            for (int r = 0; r < kAssignMaxCand; ++r) {   // sorted insert (values are >= 0)
                if (r < n_cand && v > top[r]) {
                    const float t = top[r];
                    top[r] = v;
                    v = t;
                }
            }
        }
        float sum = 0.0f;
        for (int r = 0; r < n_cand; ++r) {        // the column's n_cand largest, in descending order (torch.topk, :98)
            const uint32_t mine = top[0] < 0.0f ? 0u : __float_as_uint(top[0]);
            const uint32_t m = __reduce_max_sync(0xffffffffu, mine);
            const unsigned who = __ballot_sync(0xffffffffu, mine == m && top[0] >= 0.0f);
            if (who != 0u && lane == __ffs(who) - 1) {
#pragma unroll
                for (int q = 0; q + 1 < kAssignMaxCand; ++q) top[q] = top[q + 1];
                top[kAssignMaxCand - 1] = -1.0f;
            }
            sum = __fadd_rn(sum, __uint_as_float(m));   // topk_ious.sum(0), :101
        }
        int k = (int)sum;                          // .int(): truncation
        k = max(k, min_k);                         // torch.clamp(min=...), :101
        k = min(k, num_priors);                    // (torch.topk would refuse k > num_priors)
        if (lane == 0) ks[g] = k;
    }
    __syncthreads();

    // ---- 2. column by column: the k smallest costs among the rows, taken rows count as INFINITY (:105-111) -------------------
    const int p = tid;
    const bool real = p < num_priors;
    bool taken = false;
    int cnt = 0, first = -1;
    uint32_t parity = 0u;
    for (int g = 0; g < num_gt; ++g) {
        const int k = ks[g];
        const float v = real ? (taken ? kAssignInfinity : cost[(size_t)p * num_gt + g]) : 0.0f;
        const u64 mykey = real ? (((u64)float_order_key(v) << 32) | (uint32_t)p) : kNone64;
        bool picked = false;
        for (int r = 0; r < k; ++r) {
            const u64 wm = warp_min_u64(picked ? kNone64 : mykey);
            if (lane == 0) wred[parity][warp] = wm;
            __syncthreads();
            const u64 best = warp_min_u64(lane < nwarps ? wred[parity][lane] : kNone64);
            parity ^= 1u;
            if (best == kNone64) break;
            if (mykey == best) {
                picked = true;
                ++cnt;                             // matching_matrix[pos_idx, gt_idx] = 1.0
                if (first < 0) first = g;
            }
        }
        if (picked) taken = true;                  // cost4match[pos_idx, :] = INFINITY
    }

    // ---- 3. several ground truths for one prior: the one of least cost (:114-120) ---------------------------------------------
    int gt = first;
    if (real && cnt > 1) {
        float bestc = cost[(size_t)p * num_gt];
        gt = 0;
        for (int g = 1; g < num_gt; ++g) {
            const float c = cost[(size_t)p * num_gt + g];
            if (c < bestc) { bestc = c; gt = g; }
        }
    }

    // ---- 4. matched priors in ascending order (:122-124) -----------------------------------------------------------------------
    const bool matched = real && cnt > 0;
    const unsigned bal = __ballot_sync(0xffffffffu, matched);
    if (lane == 0) wcount[warp] = __popc(bal);
    __syncthreads();
    int base = 0, total = 0;
    for (int w = 0; w < nwarps; ++w) {
        const int c = wcount[w];
        if (w < warp) base += c;
        total += c;
    }
    if (matched) {
        const int pos = base + __popc(bal & ((1u << lane) - 1u));
        prior_idx[b * (size_t)num_priors + pos] = p;
        gt_idx[b * (size_t)num_priors + pos] = gt;
    }
    if (tid == 0) count[b] = total;
}

}  // namespace phnms

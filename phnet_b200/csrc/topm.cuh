// topm.cuh -- kernel (1) of the fast path: per-frame selection of the kTopM best-ranked proposals.
//
// Replaces, for the first greedy rounds of a frame, the ordering step `scores.sort(0, true)` of
// libs/ops/csrc/nms.cpp:51: the fused kernel only ever needs the FIRST few entries of that order (one per kept lane,
// top_k = 4 or 8 in PHNet), so a warp-level select of the kTopM smallest rank keys replaces the sort.  One warp per
// frame; keys are the radix-twiddled scores of common.cuh (ties broken by index, exactly the stable order), or, for
// frames of <= 32 proposals under the torch sort model, the positions produced by ATen's bitonic network.
// Output: per frame a block of kTopM candidate slots (32-byte header + row), in rank order (see the end of the kernel).
#pragma once
#include "common.cuh"
#include "fused_reg.cuh"

namespace phnms {

constexpr int kTopmWarps = 4;

__global__ void __launch_bounds__(kTopmWarps * 32) phnms_topm_kernel(const float *__restrict__ props,
                                                                    const float *__restrict__ scores,
                                                                    const int32_t *__restrict__ n_valid, long long F,
                                                                    int N, int n_off, int sort_model,
                                                                    int *__restrict__ topm) {
    __shared__ float bit_key[kTopmWarps][32];
    __shared__ int bit_val[kTopmWarps][32];
    __shared__ int bit_ok[kTopmWarps][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long f = (long long)blockIdx.x * kTopmWarps + warp;
    if (f >= F) return;
    int n = N;
    if (n_valid) n = max(0, min(n_valid[f], N));
    const float *sc = scores + (size_t)f * N;
    u64 mine = kNone64;  // lane j ends up with the rank key (key << 32 | index) of the j-th ranked proposal

    if (sort_model == 0 && n <= 32 && n >= 2) {  // ATen bitonicSortKVInPlace<block_dim_x = 16> (SortUtils.cuh:45-163)
        float *bk = bit_key[warp];
        int *bv = bit_val[warp], *bo = bit_ok[warp];
        bo[lane] = lane < n;
        bk[lane] = lane < n ? sc[lane] : 0.0f;
        bv[lane] = lane < n ? lane : 0;
        __syncwarp();
        for (unsigned size = 2; size <= 32; size *= 2) {
            const bool flag = (size != 32) && ((lane & (size / 2)) != 0);
            for (unsigned stride = size / 2; stride > 0; stride /= 2) {
                if (lane < 16) {
                    const unsigned pa = 2 * lane - (lane & (stride - 1)), pb = pa + stride;
                    const float ka = bk[pa], kb = bk[pb];
                    const int oa = bo[pa], ob = bo[pb];
                    const bool sw = (gt_nan(ka, kb) && oa) || !ob;
                    if (sw == flag) {
                        const int va = bv[pa], vb = bv[pb];
                        bk[pa] = kb; bk[pb] = ka;
                        bv[pa] = vb; bv[pb] = va;
                        bo[pa] = ob; bo[pb] = oa;
                    }
                }
                __syncwarp();
            }
        }
        if (lane < n) mine = ((u64)(uint32_t)lane << 32) | (uint32_t)bv[lane];  // the sorted position is the rank key
    } else {
        const bool nan_first = sort_model == 1;
        constexpr int kRegs = 32;  // frames of up to 1024 proposals keep their keys in registers
        uint32_t kreg[kRegs];
        const bool in_regs = n <= 32 * kRegs;
        if (in_regs) {
#pragma unroll
            for (int q = 0; q < kRegs; ++q) {
                const int i = lane + 32 * q;
                kreg[q] = i < n ? key_desc(sc[i], nan_first) : 0xffffffffu;
            }
        }
        u64 prev = 0;
        for (int j = 0; j < kTopM; ++j) {
            u64 best = kNone64;
            if (in_regs) {
#pragma unroll
                for (int q = 0; q < kRegs; ++q) {
                    const int i = lane + 32 * q;
                    const u64 K = ((u64)kreg[q] << 32) | (uint32_t)i;
                    if (i < n && (j == 0 || K > prev) && K < best) best = K;
                }
            } else {
                for (int i = lane; i < n; i += 32) {
                    const u64 K = ((u64)key_desc(sc[i], nan_first) << 32) | (uint32_t)i;
                    if ((j == 0 || K > prev) && K < best) best = K;
                }
            }
            best = warp_min_u64(best);
            if (best == kNone64) break;
            prev = best;
            if (lane == j) mine = best;
        }
    }

    // slot j of the frame's candidate block = {key, index, start, end, mask0, mask1, mask2, aux} + the proposal's row
    // padded to a multiple of 4 words -- byte for byte what the fused kernel keeps in shared memory, so that it can pull
    // the whole block with one bulk copy.  aux of slot 0 = number of valid candidates.
    const int P = 5 + n_off, P4 = (P + 3) & ~3, slot_words = 8 + P4;
    int *blk = topm + (size_t)f * kTopM * slot_words;
    const bool ok = lane < kTopM && mine != kNone64;
    const int count = __popc(__ballot_sync(0xffffffffu, ok));
    if (lane < kTopM) {
        uint4 h0 = make_uint4(0xffffffffu, 0xffffffffu, 0u, 0u), h1 = make_uint4(0u, 0u, 0u, (uint32_t)count);
        if (ok) {
            const uint32_t idx = (uint32_t)mine;
            const float *row = props + ((size_t)f * N + idx) * P;
            const int st = lane_start(row[2], n_off);
            const int en = lane_end(row[4], st, n_off);
            uint32_t m[3];
            range_mask<3>(st, en, m);
            h0 = make_uint4((uint32_t)(mine >> 32), idx, (uint32_t)st, (uint32_t)en);
            h1 = make_uint4(m[0], m[1], m[2], (uint32_t)count);
        }
        uint4 *out = reinterpret_cast<uint4 *>(blk + (size_t)lane * slot_words);
        out[0] = h0;
        out[1] = h1;
    }
    for (int j = 0; j < kTopM; ++j) {   // rows: the warp copies row j with coalesced loads and stores
        const u64 kj = __shfl_sync(0xffffffffu, mine, j);
        float *dst = reinterpret_cast<float *>(blk + (size_t)j * slot_words + 8);
        if (kj == kNone64) {
            for (int i = lane; i < P4; i += 32) dst[i] = 0.0f;
        } else {
            const float *row = props + ((size_t)f * N + (uint32_t)kj) * P;
            for (int i = lane; i < P4; i += 32) dst[i] = i < P ? row[i] : 0.0f;
        }
    }
}

}  // namespace phnms

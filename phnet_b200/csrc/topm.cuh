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

// dynamic shared memory: per warp 33 * ceil(N / 32) u32 keys, stored group-major ([i % 32][i / 32], row pitch G+1... see
// kidx) so that both the owner lane's column walk and the warp's row walk of one group are bank-conflict free.
__host__ __device__ inline size_t topm_keys_bytes(int N, int warps) { return (size_t)warps * 32 * (((N + 31) / 32) | 1) * 4; }
// + per warp: kTopM candidate rows (padded to 96 words), their bounds and one adjacency word each
constexpr int kTopmRowWords = 97;   // odd pitch: candidate rows start in different banks
inline size_t topm_smem_bytes(int N, int warps) {
    return topm_keys_bytes(N, warps) + (size_t)warps * kTopM * (kTopmRowWords + 2 + 1) * 4;
}

__global__ void __launch_bounds__(kTopmWarps * 32) phnms_topm_kernel(const float *__restrict__ props,
                                                                    const float *__restrict__ scores,
                                                                    const int32_t *__restrict__ n_valid, long long F,
                                                                    int N, int n_off, int sort_model, int count, float thr,
                                                                    int *__restrict__ topm,
                                                                    unsigned long long *__restrict__ claim_ctr,
                                                                    long long top_k, long long *__restrict__ keep) {
    extern __shared__ __align__(16) unsigned char smem_topm[];
    __shared__ float bit_key[kTopmWarps][32];
    __shared__ int bit_val[kTopmWarps][32];
    __shared__ int bit_ok[kTopmWarps][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (blockIdx.x == 0 && threadIdx.x == 0) *claim_ctr = 0ull;   // the fused kernel that follows hands out frames from 0
    const long long f = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
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
        // Warp-level select of the kTopM smallest rank keys.  Lane l owns "group l" = proposals l, l+32, l+64, ...;
        // it keeps the group's smallest not-yet-picked key in a register.  Per pick: one warp arg-min over the 32 group
        // minima, then the whole warp rescans the winning group (one key per lane) for its next minimum.
        const bool nan_first = sort_model == 1;
        const int G = (N + 31) / 32, pitch = G | 1;   // odd pitch: conflict-free both ways
        uint32_t *kb = reinterpret_cast<uint32_t *>(smem_topm) + (size_t)warp * 32 * pitch;
        u64 gmin = kNone64;
        if (G <= 32) {   // up to 1024 proposals: all score loads are issued before the first one is consumed
            float sv[32];
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                const int i = lane + 32 * q;
                sv[q] = (q < G && i < n) ? sc[i] : 0.0f;
            }
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                const int i = lane + 32 * q;
                if (q < G) {
                    uint32_t k = 0xffffffffu;
                    if (i < n) {
                        k = key_desc(sv[q], nan_first);
                        gmin = min(gmin, ((u64)k << 32) | (uint32_t)i);
                    }
                    kb[lane * pitch + q] = k;
                }
            }
        } else {
            for (int q = 0; q < G; ++q) {
                const int i = lane + 32 * q;
                uint32_t k = 0xffffffffu;
                if (i < n) {
                    k = key_desc(sc[i], nan_first);
                    gmin = min(gmin, ((u64)k << 32) | (uint32_t)i);
                }
                kb[lane * pitch + q] = k;
            }
        }
        __syncwarp();
        for (int j = 0; j < count; ++j) {
            const u64 best = warp_min_u64(gmin);
            if (best == kNone64) break;
            if (lane == j) mine = best;
            const int g = (int)((uint32_t)best & 31u);       // the group (== lane) that owned the pick
            u64 cand = kNone64;
            for (int q = lane; q < G; q += 32) {              // rescan group g: proposal g + 32 q
                const int i = g + 32 * q;
                const u64 K = ((u64)kb[g * pitch + q] << 32) | (uint32_t)i;
                if (i < n && K > best) cand = min(cand, K);
            }
            cand = warp_min_u64(cand);
            if (lane == g) gmin = cand;
        }
    }

    // slot j of the frame's candidate block = {key, index, start, end, mask0, mask1, mask2, aux} + the proposal's row
    // padded to a multiple of 4 words -- byte for byte what the fused kernel keeps in shared memory, so that it can pull
    // the whole block with one bulk copy.  aux = adjacency bits << 16 | number of valid candidates (slot 1: | the kept set
    // of the planned greedy scan instead of the count).
    const int P = 5 + n_off, P4 = (P + 3) & ~3, slot_words = 8 + P4;
    int *blk = topm + (size_t)f * count * slot_words;
    const bool ok = lane < count && mine != kNone64;
    const int nfound = __popc(__ballot_sync(0xffffffffu, ok));
    uint4 h0 = make_uint4(0xffffffffu, 0xffffffffu, 0u, 0u), h1 = make_uint4(0u, 0u, 0u, (uint32_t)nfound);
    if (ok) {
        const uint32_t idx = (uint32_t)mine;
        const float *row = props + ((size_t)f * N + idx) * P;
        const int st = lane_start(row[2], n_off);
        const int en = lane_end(row[4], st, n_off);
        uint32_t m[3];
        range_mask<3>(st, en, m);
        h0 = make_uint4((uint32_t)(mine >> 32), idx, (uint32_t)st, (uint32_t)en);
        h1 = make_uint4(m[0], m[1], m[2], (uint32_t)nfound);
    }
    // rows: the warp copies row j with coalesced loads and stores; all loads are issued before the first store
    if (P4 <= 96) {
        float rv[kTopM][3];
#pragma unroll
        for (int j = 0; j < kTopM; ++j) {
            if (j >= count) break;
            const u64 kj = __shfl_sync(0xffffffffu, mine, j);
            const float *row = props + ((size_t)f * N + (kj == kNone64 ? 0u : (uint32_t)kj)) * P;
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                const int i = lane + 32 * t;
                rv[j][t] = (kj != kNone64 && i < P) ? row[i] : 0.0f;
            }
        }
        // candidate-vs-candidate predicate: bit j of adj[i] = devIoU(candidate i, candidate j), j ranked after i.  With it
        // the fused kernel knows, before touching a single column, which candidates the greedy scan keeps.
        float *crow = reinterpret_cast<float *>(smem_topm + topm_keys_bytes(N, blockDim.x >> 5)) +
                      (size_t)warp * kTopM * (kTopmRowWords + 3);
        int *cse = reinterpret_cast<int *>(crow + kTopM * kTopmRowWords);
        uint32_t *adj = reinterpret_cast<uint32_t *>(cse + 2 * kTopM);
#pragma unroll
        for (int j = 0; j < kTopM; ++j) {
            if (j >= count) break;
#pragma unroll
            for (int t = 0; t < 3; ++t) crow[j * kTopmRowWords + lane + 32 * t] = rv[j][t];
        }
        if (lane < kTopM) {
            cse[2 * lane] = (int)h0.z;
            cse[2 * lane + 1] = (int)h0.w;
            adj[lane] = 0u;
        }
        __syncwarp();
        for (int pr = lane; pr < nfound * (nfound - 1) / 2; pr += 32) {
            int i = 0, rem = pr;
            while (rem >= nfound - 1 - i) { rem -= nfound - 1 - i; ++i; }
            const int j = i + 1 + rem;
            if (pair_hit_scalar(crow + i * kTopmRowWords, crow + j * kTopmRowWords, cse[2 * i], cse[2 * i + 1], cse[2 * j],
                                cse[2 * j + 1], thr))
                atomicOr(&adj[i], 1u << j);
        }
        __syncwarp();
        // The greedy scan over the candidates (nms_collect, :111-136, restricted to them) is done here, once per frame:
        // candidate i is kept iff no kept candidate before it suppresses it; the scan stops after top_k kept lanes (:133).
        // Kept candidates are final -- they are the first kept lanes of the frame, in this order -- so their indices go
        // straight to keep[f, 0 ..] (:118); the fused kernel gets the kept set as a bitmask in slot 1's aux word.
        uint32_t alive = nfound >= 32 ? 0xffffffffu : ((1u << nfound) - 1u), kept = 0u;
        {
            long long nk = 0;
            for (int i = 0; i < nfound; ++i) {
                if ((alive >> i) & 1u) {
                    kept |= 1u << i;
                    alive &= ~adj[i];
                    if (lane == i && keep) keep[(size_t)f * N + nk] = (long long)(uint32_t)mine;
                    ++nk;
                    if (nk == top_k) break;
                }
            }
        }
        if (lane < count) h1.w = (adj[lane] << 16) | (lane == 1 ? kept & 0xffffu : (uint32_t)nfound);
#pragma unroll
        for (int j = 0; j < kTopM; ++j) {
            if (j >= count) break;
            float *dst = reinterpret_cast<float *>(blk + (size_t)j * slot_words + 8);
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                const int i = lane + 32 * t;
                if (i < P4) dst[i] = rv[j][t];
            }
        }
    } else {
        for (int j = 0; j < count; ++j) {
            const u64 kj = __shfl_sync(0xffffffffu, mine, j);
            float *dst = reinterpret_cast<float *>(blk + (size_t)j * slot_words + 8);
            const float *row = props + ((size_t)f * N + (kj == kNone64 ? 0u : (uint32_t)kj)) * P;
            for (int i = lane; i < P4; i += 32) dst[i] = (kj != kNone64 && i < P) ? row[i] : 0.0f;
        }
    }
    if (lane < count) {   // aux = (adjacency bits << 16) | number of valid candidates
        uint4 *out = reinterpret_cast<uint4 *>(blk + (size_t)lane * slot_words);
        out[0] = h0;
        out[1] = h1;
    }
}

}  // namespace phnms

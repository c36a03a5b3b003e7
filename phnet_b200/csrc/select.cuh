// select.cuh -- kernel (1) of the streaming fast path: the greedy scan itself, run on as few proposals as it needs.
//
// What it replaces: `scores.sort(0, true)` (libs/ops/csrc/nms.cpp:51) and the serial scan `nms_collect`
// (libs/ops/csrc/nms_kernel.cu:99-143).  nms_collect walks the sorted order, keeps a proposal iff no earlier KEPT proposal
// covers it, and stops after top_k kept lanes (:133).  So only a prefix of the order is ever looked at, and a proposal of
// that prefix only has to be compared with the lanes kept before it.  One warp per frame does exactly that:
//
//   repeat:  draw the next kSelBatch proposals in rank order (warp-level select over radix-twiddled keys, ties by index ==
//            the stable order; <= 32 proposals under the torch sort model: ATen's bitonic network replayed),
//            fetch their rows, evaluate devIoU (:26-48) of each against the lanes kept so far, then among the survivors,
//            run the scan over the batch (a survivor is kept iff no kept survivor before it covers it)
//   until    top_k lanes are kept, or every proposal of the frame was drawn, or `cap` proposals were drawn.
//
// Output per frame: a block {nk, open, n} + nk slots {rank key, index, start, end, in-range masks, row} -- byte for byte what
// phnms_stream_kernel (stream.cuh) evaluates against every proposal of the frame -- plus keep[f, 0..nk) and num_keep[f].
// A frame is left OPEN when the cap was hit with fewer than top_k lanes kept and proposals left undrawn: if the streaming
// pass then finds a proposal that no kept lane covers, the frame goes on the resume list (stream.cuh) and is redone by the
// cluster kernel; if everything is covered (a road with fewer lanes than top_k) the frame is complete as it stands.
// The cost follows the input: 8 draws for a frame whose first candidates are distinct lanes, up to `cap` draws otherwise.
#pragma once
#include "common.cuh"
#include "fused_reg.cuh"

namespace phnms {

#ifndef PHNMS_SELECT_CHUNK
#define PHNMS_SELECT_CHUNK 16
#endif
#ifndef PHNMS_SELECT_CTAS
#define PHNMS_SELECT_CTAS 6
#endif
constexpr int kSelWarps = 4;
constexpr int kSelBatch = 8;       // proposals drawn per batch (one per lane 0..7)
constexpr int kSelCapDefault = 64; // draws per frame before it is left open
constexpr int kStreamMaxK = 8;     // kept lanes the streaming path carries per frame (PHNet: max_lanes 4, VIL-100 8)
constexpr int kBlkHdr = 32;        // bytes: {nk, open, n, 0, 0, 0, 0, 0}

struct SelectParams {
    const float *props;
    const float *scores;
    const int32_t *n_valid;
    long long F;
    int N, n_off, sort_model, top_k, cap;
    float thr;
    unsigned char *blocks;   // [F] blocks of block_bytes = kBlkHdr + top_k * (kHdr + 4 * round4(5 + n_off))
    int block_bytes;
    int *flags;              // [F] "already on the resume list" (cleared here for open frames)
    unsigned int *ctrs;      // ctrs[0] = length of the resume list (cleared here)
    long long *keep;
    long long *num_keep;
};

__host__ __device__ inline int select_warp_words(int N, int n_off, int top_k) {
    const int P4 = (5 + n_off + 3) & ~3;
    const int w = top_k * (8 + P4) + kSelBatch * (P4 | 1) + 2 * kSelBatch + kSelBatch + 32 * (((N + 31) / 32) | 1);
    return (w + 3) & ~3;
}
inline size_t select_smem_bytes(int N, int n_off, int top_k, int warps) {
    return (size_t)warps * select_warp_words(N, n_off, top_k) * 4;
}

// key_desc (common.cuh) without the NaN-first case, in four integer instructions: ascending-u32 == descending score, -0.0 folded
// onto +0.0.  Positive floats: the low 31 bits inverted (larger score -> smaller key, top bit 0); negative floats: the bit
// pattern itself (top bit 1, larger magnitude -> larger key).  Same values as key_desc(s, false).
__device__ __forceinline__ uint32_t key_desc_fast(float s) {
    uint32_t u = __float_as_uint(s);
    if (u == 0x80000000u) u = 0u;
    return u ^ ((uint32_t)((int32_t)~u >> 31) & 0x7fffffffu);
}

// the 28 pairs (a < b) of a batch of 8, packed 4 bits per pair: lane -> (a, b)
__device__ __forceinline__ void batch_pair(int lane, int &a, int &b) {
    // (0,1)(0,2)(0,3)(0,4)(0,5)(0,6)(0,7)(1,2)(1,3)(1,4)(1,5)(1,6)(1,7)(2,3)(2,4)(2,5) | (2,6)(2,7)(3,4)(3,5)(3,6)(3,7)(4,5)(4,6)(4,7)(5,6)(5,7)(6,7)
    const unsigned long long A0 = 0x2221111110000000ull, B0 = 0x5437654327654321ull;
    const unsigned long long A1 = 0x0000655444333322ull, B1 = 0x0000776765765476ull;
    const int sh = 4 * (lane & 15);
    a = (int)(((lane < 16 ? A0 : A1) >> sh) & 15ull);
    b = (int)(((lane < 16 ? B0 : B1) >> sh) & 15ull);
}

// ---- the two steps of a batch, shared by phnms_select_kernel and the fused get_lanes kernel (frontend.cuh) -------------------
// (Measured and rejected: caching each group's SECOND smallest key in a register so that most draws need no rescan -- the extra
// compare / select per key in the key pass costs more than the rescans save: select 80 -> 92 us at the headline shape.)
// Draw: the next (up to) kSelBatch proposals in rank order.  Lane l owns "group l" = proposals l, l + 32, ... and keeps the
// group's smallest not-yet-drawn rank key in `gmin`; per pick one warp arg-min over the 32 group minima, then the warp
// rescans the winning group (one key per lane) for its next minimum.  `valid(i, q)`: proposal i = g + 32 q exists.
template <typename ValidFn>
__device__ __forceinline__ int select_draw_batch(u64 &gmin, const uint32_t *kb, int pitch, int G, int lane, u64 &myc, ValidFn valid) {
    int nb = 0;
    myc = kNone64;
    for (int j = 0; j < kSelBatch; ++j) {
        const u64 best = warp_min_u64(gmin);
        if (best == kNone64) break;
        if (lane == j) myc = best;
        ++nb;
        const int g = (int)((uint32_t)best & 31u);       // the group (== lane) that owned the pick
        u64 cand = kNone64;
        for (int q = lane; q < G; q += 32) {              // rescan group g: proposal g + 32 q
            const int i = g + 32 * q;
            const u64 K = ((u64)kb[g * pitch + q] << 32) | (uint32_t)i;
            if (valid(i, q) && K > best) cand = min(cand, K);
        }
        cand = warp_min_u64(cand);
        if (lane == g) gmin = cand;
    }
    return nb;
}

// Scan: candidates j < nb (rank key of candidate j in lane j's `myc`, rows in brow) against the lanes kept so far (the mask
// rows nms_collect would OR into remv, nms_kernel.cu:116-122), then the survivors against each other (strict upper triangle
// in rank order, :85-87) and the greedy scan over the batch (:111-136).  Newly kept lanes are appended to `slots` (header +
// row, the layout the streaming kernel reads); on_keep(k, Ka) is called by every lane for the k-th kept lane.  Returns nk.
template <typename OnKeep>
__device__ __forceinline__ int select_scan_batch(int nb, u64 myc, int nk, int top_k, int n_off, float thr, uint32_t *slots,
                                                 float *brow, int *bse, uint32_t *adj, int lane, OnKeep on_keep) {
    const int P4 = (5 + n_off + 3) & ~3, slot_words = 8 + P4, bp = P4 | 1;
    if (lane < nb) {
        const int st = lane_start(brow[lane * bp + 2], n_off);       // nms_kernel.cu:29-30
        bse[2 * lane] = st;
        bse[2 * lane + 1] = lane_end(brow[lane * bp + 4], st, n_off); // :32-34
    }
    if (lane < kSelBatch) adj[lane] = 0u;
    __syncwarp();
    uint32_t supp = 0u;
    const int tot = nk * nb;
    for (int b0 = 0; b0 < tot; b0 += 32) {
        const int pr = b0 + lane;
        uint32_t bit = 0u;
        if (pr < tot) {
            const int i = pr / nb, j = pr - i * nb;
            const uint32_t *sl = slots + i * slot_words;
            if (pair_hit_scalar(reinterpret_cast<const float *>(sl + 8), brow + j * bp, (int)sl[2], (int)sl[3], bse[2 * j],
                                bse[2 * j + 1], thr))
                bit = 1u << j;
        }
        supp |= __reduce_or_sync(0xffffffffu, bit);
    }
    const uint32_t surv = ~supp & ((1u << nb) - 1u);
    const int s = __popc(surv), need = top_k - nk;
    if (s >= 2 && need >= 2) {
        int a, b;
        batch_pair(lane, a, b);
        if (lane < 28 && ((surv >> a) & (surv >> b) & 1u)) {
            if (pair_hit_scalar(brow + a * bp, brow + b * bp, bse[2 * a], bse[2 * a + 1], bse[2 * b], bse[2 * b + 1], thr))
                atomicOr(&adj[a], 1u << b);
        }
        __syncwarp();
    }
    uint32_t alive = surv;
    while (alive && nk < top_k) {
        const int a = __ffs(alive) - 1;
        alive &= ~(1u << a);
        alive &= ~adj[a];
        const u64 Ka = __shfl_sync(0xffffffffu, myc, a);
        uint32_t *sl = slots + nk * slot_words;
        for (int i = lane; i < P4; i += 32) sl[8 + i] = __float_as_uint(brow[a * bp + i]);
        if (lane == 0) {
            const int st = bse[2 * a], en = bse[2 * a + 1];
            uint32_t m[3];
            range_mask<3>(st, en, m);
            sl[0] = (uint32_t)(Ka >> 32); sl[1] = (uint32_t)Ka; sl[2] = (uint32_t)st; sl[3] = (uint32_t)en;
            sl[4] = m[0]; sl[5] = m[1]; sl[6] = m[2]; sl[7] = 0u;
        }
        on_keep(nk, Ka);
        ++nk;
    }
    __syncwarp();
    return nk;
}

// REC: also store every frame's compact record {keep[0 .. top_k), num} to every destination of `rec` (phnms_forward_collect_f32).  A
// template parameter and a separate argument: carrying the descriptor inside SelectParams cost the plain kernel 7 % (80 -> 86 us).
template <bool REC>
__global__ void __launch_bounds__(kSelWarps * 32, PHNMS_SELECT_CTAS) phnms_select_kernel(const SelectParams sp, const RecordSink rec) {
    extern __shared__ __align__(16) unsigned char smem_sel[];
    __shared__ float bit_key[kSelWarps][32];
    __shared__ int bit_val[kSelWarps][32];
    __shared__ int bit_ok[kSelWarps][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (blockIdx.x == 0 && threadIdx.x == 0) sp.ctrs[0] = 0u;   // the resume list of this call starts empty
    const long long f = (long long)blockIdx.x * nw + warp;
    if (f >= sp.F) return;
    const int N = sp.N, n_off = sp.n_off, P = 5 + n_off, P4 = (P + 3) & ~3, slot_words = 8 + P4, bp = P4 | 1;
    const int top_k = sp.top_k;
    int n = N;
    if (sp.n_valid) n = max(0, min(sp.n_valid[f], N));
    const int G = (N + 31) / 32, pitch = G | 1;   // odd pitch: conflict-free both ways

    uint32_t *slots = reinterpret_cast<uint32_t *>(smem_sel) + (size_t)warp * select_warp_words(N, n_off, top_k);
    float *brow = reinterpret_cast<float *>(slots + top_k * slot_words);   // [kSelBatch][bp] rows of the current batch
    int *bse = reinterpret_cast<int *>(brow + kSelBatch * bp);              // [kSelBatch][2] their (start, end)
    uint32_t *adj = reinterpret_cast<uint32_t *>(bse + 2 * kSelBatch);      // [kSelBatch] survivor-vs-survivor hits
    uint32_t *kb = adj + kSelBatch;                                         // [32][pitch] rank keys, group-major

    const float *sc = sp.scores + (size_t)f * N;
    const bool bitonic = sp.sort_model == 0 && n <= 32 && n >= 2;
    u64 sorted = kNone64;   // bitonic: lane j holds the j-th ranked proposal
    u64 gmin = kNone64;     // select: smallest not-yet-drawn rank key of this lane's group (proposals lane, lane+32, ...)

    if (bitonic) {          // ATen bitonicSortKVInPlace<block_dim_x = 16> (SortUtils.cuh:45-163), see topm.cuh
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
        if (lane < n) sorted = ((u64)(uint32_t)lane << 32) | (uint32_t)bv[lane];  // the sorted position is the rank key
    } else if (n > 0) {
        // Rank keys of the whole frame -> shared memory (lane l owns proposals l, l + 32, ...), the lane's smallest key in a
        // register.  Chunks of 8 keys per lane; a chunk that lies entirely inside the frame needs no bounds predicates, and
        // within a lane a strict `<` keeps the earliest (smallest-index) proposal among equal keys.
        const bool nan_first = sp.sort_model == 1;
        uint32_t *kbl = kb + lane * pitch;
        const int cnt = lane < n ? ((n - lane + 31) >> 5) : 0;   // keys this lane owns
        uint32_t bestk = 0xffffffffu;
        int bestq = 0;
        constexpr int KC = PHNMS_SELECT_CHUNK;   // keys per lane whose loads are in flight together
        for (int q0 = 0; q0 < G; q0 += KC) {
            if (!nan_first && 32 * (q0 + KC) <= n) {
                float sv[KC];
#pragma unroll
                for (int u = 0; u < KC; ++u) sv[u] = sc[lane + 32 * (q0 + u)];
#pragma unroll
                for (int u = 0; u < KC; ++u) {
                    const uint32_t k = key_desc_fast(sv[u]);
                    if (k < bestk) { bestk = k; bestq = q0 + u; }
                    kbl[q0 + u] = k;
                }
            } else {
                float sv[KC];
#pragma unroll
                for (int u = 0; u < KC; ++u) sv[u] = (q0 + u < cnt) ? sc[lane + 32 * (q0 + u)] : 0.0f;
#pragma unroll
                for (int u = 0; u < KC; ++u) {
                    const int q = q0 + u;
                    uint32_t k = 0xffffffffu;
                    if (q < cnt) {
                        k = key_desc(sv[u], nan_first);
                        if (k < bestk) { bestk = k; bestq = q; }   // (all-ones keys: bestq stays 0, the earliest)
                    }
                    if (q < G) kbl[q] = k;
                }
            }
        }
        if (cnt > 0) gmin = ((u64)bestk << 32) | (uint32_t)(lane + 32 * bestq);
        __syncwarp();
    }

    int nk = 0, drawn = 0;
    bool open = false;
    while (n > 0) {
        // ---- draw the next batch in rank order: lane j < nb holds candidate j ------------------------------------
        u64 myc = kNone64;
        int nb = 0;
        if (bitonic) {
            nb = min(kSelBatch, n - drawn);
            const u64 v = __shfl_sync(0xffffffffu, sorted, (drawn + lane) & 31);
            if (lane < nb) myc = v;
        } else {
            nb = select_draw_batch(gmin, kb, pitch, G, lane, myc, [&](int i, int) { return i < n; });
        }
        if (nb == 0) break;
        // ---- their rows: coalesced loads, all issued before the first store ---------------------------------------
        {
            float rv[kSelBatch][3];
#pragma unroll
            for (int j = 0; j < kSelBatch; ++j) {
                const u64 kj = __shfl_sync(0xffffffffu, myc, j);
                const float *row = sp.props + ((size_t)f * N + (j < nb ? (uint32_t)kj : 0u)) * P;
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    const int i = lane + 32 * t;
                    rv[j][t] = (j < nb && i < P) ? row[i] : 0.0f;
                }
            }
#pragma unroll
            for (int j = 0; j < kSelBatch; ++j)
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    const int i = lane + 32 * t;
                    if (j < nb && i < P4) brow[j * bp + i] = rv[j][t];
                }
        }
        __syncwarp();
        nk = select_scan_batch(nb, myc, nk, top_k, n_off, sp.thr, slots, brow, bse, adj, lane, [&](int k, u64 Ka) {
            if (lane == 0) sp.keep[(size_t)f * N + k] = (long long)(uint32_t)Ka;   // :118
        });
        drawn += nb;
        if (nk == top_k || drawn >= n) break;   // :133 / every proposal of the frame was drawn
        if (drawn >= sp.cap) {
            open = true;
            break;
        }
    }

    // ---- the frame's block: header + nk slots, 16-byte stores --------------------------------------------------------------
    unsigned char *blk = sp.blocks + (size_t)f * sp.block_bytes;
    if (lane < 2) {
        const uint4 h = lane == 0 ? make_uint4((uint32_t)nk, open ? 1u : 0u, (uint32_t)n, 0u) : make_uint4(0u, 0u, 0u, 0u);
        reinterpret_cast<uint4 *>(blk)[lane] = h;
    }
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(slots);
        uint4 *dst = reinterpret_cast<uint4 *>(blk + kBlkHdr);
        for (int i = lane; i < nk * slot_words / 4; i += 32) dst[i] = src[i];
    }
    if (lane == 0) {
        sp.num_keep[f] = (long long)nk;   // :142 (nk <= top_k)
        if (open) sp.flags[f] = 0;
    }
    // (an open frame that the resume pass redoes gets its record rewritten there)
    if (REC && lane < rec.width)
        record_store(rec, f, lane, lane == rec.width - 1 ? (long long)nk : (lane < nk ? (long long)slots[lane * slot_words + 1] : 0ll));
}

}  // namespace phnms

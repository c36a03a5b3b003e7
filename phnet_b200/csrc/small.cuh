// small.cuh -- the whole op for a SMALL call in one launch: one CTA per frame, every proposal of the frame in registers.
//
// PHNet calls the op once per frame (libs/models/Router4OL.py:460-465: `keep, num_to_keep, _ = nms(...)` on the <= 240 priors
// that passed the confidence filter), so what matters there is the latency of one launch, not bandwidth.  This kernel is the
// shortest dependent chain I could find for the reference's three steps (nms.cpp:51 sort, nms_kernel.cu:50-96 mask rows,
// :99-143 collect) on one frame of <= 512 proposals:
//
//   1. every warp fetches its own 32 rows with ONE bulk copy (UBLKCP; <= 3 unaligned words at either end and the scores by
//      4-byte cp.async) onto its own mbarrier -- no CTA-wide step -- and moves them to registers (one proposal per thread);
//   2. greedy rounds, one per kept lane: CTA-wide arg-min of the rank keys of the proposals still alive (REDUX in the warp,
//      one shared-memory hop across warps), the winner publishes its row in the slot format of the streaming path, all
//      threads evaluate devIoU against it over their registers (stream_eval: the reference's sequential fp32 sum) and stamp
//      `parent` -- exactly the rows of the mask nms_collect reads (:116-129), in the order it reads them;
//   3. outputs once: keep[0..nk) as the rounds go, zero padding, parent, num_to_keep = min(top_k, nk) (:139-142).
//
// Two __syncthreads per round, everything double buffered; top_k = 0 (never stops, reports 0) simply runs until no proposal
// is alive.  Frames of <= 32 proposals under the torch sort model are ordered by ATen's bitonic network (replayed, see
// select.cuh).  Used by the planner for calls of at most 2048 proposals in total (phnms.cu, make_plan).
#pragma once
#include "common.cuh"
#include "fused_reg.cuh"
#include "stream.cuh"

namespace phnms {

constexpr int kSmallMaxN = 512;

struct SmallParams {
    const float *props;
    const float *scores;
    const int32_t *n_valid;
    long long *keep;
    long long *num_keep;
    long long *parent;
    long long F;
    long long top_k;
    int N, sort_model;
    float thr;
};

struct SmallLayout {
    int off_wmin, off_bit, off_pub, off_kept, off_stash, off_slots, slot_bytes, total;
};

// shared memory: [0,128) one mbarrier per warp | wmin[2][16] u64 | bitonic scratch | 2 published slots | kept indices | header stash
// (5 words per thread) | per-warp staging
__host__ __device__ inline SmallLayout small_layout(int warps, int P) {
    SmallLayout L;
    const int P4 = (P + 3) & ~3;
    int o = 128;
    L.off_wmin = o;
    o += 2 * 16 * 8;
    L.off_bit = o;
    o += 384;
    L.off_pub = o;
    o += 2 * (kHdr + 4 * P4);
    L.off_kept = o;          // indices of the lanes kept so far (read back when the frame's record is stored)
    o += kSmallMaxN * 4;
    L.off_stash = o;
    o += warps * 32 * 5 * 4;
    o = (o + 127) & ~127;
    L.off_slots = o;
    L.slot_bytes = 128 + ((16 + 32 * P * 4 + 16 + 127) & ~127);
    o += warps * L.slot_bytes;
    L.total = o;
    return L;
}

// The request of one warp's 32 rows of frame f (the same scheme as phnms_stream_kernel's: rows keep their global address modulo
// 16, the aligned body is one bulk copy, <= 3 words at either end and the scores are 4-byte cp.async, all on the warp's mbarrier).
template <int P>
__device__ __forceinline__ void small_request(const SmallParams &sp, long long f, int r0, int nrows, uint32_t slot_s, uint32_t bar,
                                              int lane) {
    const float *src = sp.props + ((size_t)f * sp.N + r0) * P;
    const uintptr_t a0 = (uintptr_t)src;
    const uint32_t bytes = (uint32_t)nrows * (P * 4);
    if (((a0 | bytes) & 15u) == 0u) {
        if (lane == 0) {
            mbar_arrive_expect_tx(bar, bytes);
            bulk_g2s(slot_s + 128u, src, bytes, bar);
        }
    } else {
        const uintptr_t b0 = (a0 + 15) & ~(uintptr_t)15, e0 = (a0 + bytes) & ~(uintptr_t)15;
        const uint32_t D = slot_s + 128u + (uint32_t)(a0 & 15);
        if (lane == 0) {
            mbar_arrive_expect_tx(bar, (uint32_t)(e0 - b0));
            bulk_g2s(D + (uint32_t)(b0 - a0), reinterpret_cast<const void *>(b0), (uint32_t)(e0 - b0), bar);
        }
        const int hw = (int)((b0 - a0) >> 2), tw = (int)((a0 + bytes - e0) >> 2), t0 = (int)((e0 - a0) >> 2);
        if (lane >= 1 && lane - 1 < hw) cp_async_4(D + 4u * (uint32_t)(lane - 1), src + (lane - 1));
        if (lane >= 4 && lane - 4 < tw) cp_async_4(D + 4u * (uint32_t)(t0 + lane - 4), src + t0 + (lane - 4));
    }
    if (lane < nrows) cp_async_4(slot_s + 4u * (uint32_t)lane, sp.scores + (size_t)f * sp.N + r0 + lane);
    cp_async_mbar_arrive_noinc(bar);
}

// MAXT / MINB: launch bounds (frames of <= 256 proposals at 36 offsets fit three CTAs per SM in 80 registers).  (Measured and
// rejected: two proposals per thread at 36 offsets -- four warps per 240-proposal frame, two chains per thread: 0.344 vs 0.354 of the
// roofline at top_k 8, 0.570 vs 0.583 at top_k 4.)
// Persistent: CTA b takes frames b, b + gridDim.x, ...; a warp requests its rows of the NEXT frame as soon as the current ones are
// in registers, so the load of a frame overlaps the greedy rounds of the one before.
// REC: also store every frame's compact record (`rec`: {keep[0 .. top_k), num} to every destination) -- a separate instantiation, the plain one carries none of it (the kernel
// sits at its register limit: the branches alone cost 2 %).
template <int NOFF, int MAXT, int MINB, bool REC>
__global__ void __launch_bounds__(MAXT, MINB) phnms_small_kernel(const SmallParams sp, const RecordSink rec) {
    constexpr int P = 5 + NOFF, MW = (P + 31) / 32, P4 = (P + 3) & ~3, SLOT = kHdr + 4 * P4;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const SmallLayout L = small_layout(nwarps, P);
    const uint32_t smem_s = smem_u32(smem);
    const uint32_t bar = smem_s + 8u * warp;
    const uint32_t slot_s = smem_s + (uint32_t)L.off_slots + (uint32_t)warp * (uint32_t)L.slot_bytes;
    u64 *wmin = reinterpret_cast<u64 *>(smem + L.off_wmin);
    uint32_t *stash = reinterpret_cast<uint32_t *>(smem + L.off_stash);
    uint32_t *kept = reinterpret_cast<uint32_t *>(smem + L.off_kept);
    const int N = sp.N, r0 = warp * 32;
    auto rows_of = [&](long long f) {
        int n = N;
        if (sp.n_valid) n = max(0, min(sp.n_valid[f], N));
        return n;
    };

    if (lane == 0) {
        mbar_init(bar, 33);   // 1 expect_tx arrive + 32 cp.async arrives
        fence_mbar_init();
    }
    __syncwarp();
    long long f = blockIdx.x;
    uint32_t rphase = 0u, parity = 0u;
    if (f < sp.F) {
        const int nr = max(0, min(rows_of(f) - r0, 32));
        if (nr > 0) small_request<P>(sp, f, r0, nr, slot_s, bar, lane);
    }

    for (; f < sp.F; f += gridDim.x) {
        const int n = rows_of(f);
        // ---- 1. this warp's rows: staging slot -> registers; then the next frame's request goes out --------------------------------
        const int nrows = max(0, min(n - r0, 32));
        float x[1][NOFF];
        int st[1] = {0}, en[1] = {-1};
        float score = 0.0f;
        const bool real[1] = {lane < nrows};
        if (nrows > 0) {
            const uintptr_t a0 = (uintptr_t)(sp.props + ((size_t)f * N + r0) * P);
            const uint32_t row = slot_s + 128u + (uint32_t)(a0 & 15) + (uint32_t)(real[0] ? lane : 0) * (P * 4);
            mbar_wait(bar, rphase);
            rphase ^= 1u;
#pragma unroll
            for (int i = 0; i < NOFF; ++i) x[0][i] = __uint_as_float(lds_u32(row + 4u * (5 + i)));
            st[0] = lane_start(__uint_as_float(lds_u32(row + 8u)), NOFF);            // nms_kernel.cu:29-30
            en[0] = lane_end(__uint_as_float(lds_u32(row + 16u)), st[0], NOFF);      // :32-34
            score = __uint_as_float(lds_u32(slot_s + 4u * (uint32_t)(real[0] ? lane : 0)));
            // the five header words stay reachable after the slot is handed to the next frame (a kept lane publishes them)
#pragma unroll
            for (int i = 0; i < 5; ++i) stash[tid * 5 + i] = lds_u32(row + 4u * i);
        } else {
#pragma unroll
            for (int i = 0; i < NOFF; ++i) x[0][i] = 0.0f;
        }
        __syncwarp();   // every lane has read its row: the slot is free
        {
            const long long fn = f + gridDim.x;
            if (fn < sp.F) {
                const int nr = max(0, min(rows_of(fn) - r0, 32));
                if (nr > 0) small_request<P>(sp, fn, r0, nr, slot_s, bar, lane);
            }
        }
        // rank key: ascending u64 (key << 32 | index) == the order of scores.sort(0, true) (nms.cpp:51)
        uint32_t key = key_desc(real[0] ? score : 0.0f, sp.sort_model == 1);
        if (sp.sort_model == 0 && n <= 32 && n >= 2 && warp == 0) {   // ATen bitonicSortKVInPlace (SortUtils.cuh:45-163): unstable
            float *bit_key = reinterpret_cast<float *>(smem + L.off_bit);
            int *bit_val = reinterpret_cast<int *>(bit_key + 32), *bit_ok = bit_val + 32;
            bit_ok[lane] = lane < n;
            bit_key[lane] = lane < n ? score : 0.0f;
            bit_val[lane] = lane < n ? lane : 0;
            __syncwarp();
            for (unsigned size = 2; size <= 32; size *= 2) {
                const bool flag = (size != 32) && ((lane & (size / 2)) != 0);
                for (unsigned stride = size / 2; stride > 0; stride /= 2) {
                    if (lane < 16) {
                        const unsigned pa = 2 * lane - (lane & (stride - 1)), pb = pa + stride;
                        const float ka = bit_key[pa], kb = bit_key[pb];
                        const int oa = bit_ok[pa], ob = bit_ok[pb];
                        const bool sw = (gt_nan(ka, kb) && oa) || !ob;
                        if (sw == flag) {
                            const int va = bit_val[pa], vb = bit_val[pb];
                            bit_key[pa] = kb; bit_key[pb] = ka;
                            bit_val[pa] = vb; bit_val[pb] = va;
                            bit_ok[pa] = ob;  bit_ok[pb] = oa;
                        }
                    }
                    __syncwarp();
                }
            }
            int mypos = 0;
            for (int q = 0; q < 32; ++q)
                if (bit_val[q] == lane && q < n) mypos = q;
            key = (uint32_t)mypos;
            __syncwarp();
        }
        const u64 myK[1] = {real[0] ? (((u64)key << 32) | (uint32_t)(r0 + lane)) : kNone64};
        uint32_t mb[1][MW], par[1] = {0u};
        range_mask<MW>(st[0], en[0], mb[0]);

        // ---- 2. greedy rounds (nms_collect, :111-136) ---------------------------------------------------------------------------
        // Two barriers per round: rank keys across warps, then the winner's row.  (Publishing every warp's own best row
        // speculatively saves the second barrier but costs 8x the shared-memory stores: measured 0.48 vs 0.60 of the roofline.)
        // Keys and the published slot are double buffered and the parity flips after EVERY round, also the last one of a frame:
        // what a round writes was last read two rounds earlier, with a barrier in between.
        bool alive = real[0];
        int nk = 0;
        while (n > 0) {
            const u64 wm = warp_min_u64(alive ? myK[0] : kNone64);
            if (lane == 0) wmin[parity * 16 + warp] = wm;
            __syncthreads();
            const u64 best = warp_min_u64(lane < nwarps ? wmin[parity * 16 + lane] : kNone64);
            const uint32_t pub_s = smem_s + (uint32_t)L.off_pub + parity * SLOT;
            unsigned char *pub = smem + (pub_s - smem_s);
            parity ^= 1u;
            if (best == kNone64) break;   // nobody left (:116 never true again)
            if (myK[0] == best) {         // the kept lane publishes {rank key, index, start, end, in-range masks} + its row, 16 bytes at a time
                uint32_t m3[3];
                range_mask<3>(st[0], en[0], m3);
                sts_v4(pub_s, (uint32_t)(best >> 32), (uint32_t)best, (uint32_t)st[0], (uint32_t)en[0]);
                sts_v4(pub_s + 16u, m3[0], m3[1], m3[2], 0u);
                uint32_t hw[5];
#pragma unroll
                for (int i = 0; i < 5; ++i) hw[i] = stash[tid * 5 + i];
#pragma unroll
                for (int g = 0; g < P4 / 4; ++g) {
                    uint32_t w[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = 4 * g + u;
                        w[u] = i < 5 ? hw[i < 5 ? i : 0] : (i < P ? __float_as_uint(x[0][i < P && i >= 5 ? i - 5 : 0]) : 0u);
                    }
                    sts_v4(pub_s + kHdr + 16u * g, w[0], w[1], w[2], w[3]);
                }
                sp.keep[(size_t)f * N + nk] = (long long)(uint32_t)best;   // :118
                if (REC) kept[nk] = (uint32_t)best;   // (the record goes out at the end of the frame, off the rounds' critical path)
            }
            __syncthreads();
            if (!stream_eval<NOFF, 1, 1>(pub_s, 1, real, myK, st, en, mb, x, sp.thr, par, nk)) {
                // a pair with a negative common start somewhere in the warp (header words / the wrapped unsigned-char counter, :38)
                FusedParams fp;
                fp.thr = sp.thr;
                auto my_hdr = [&](int) { return sp.props + ((size_t)f * N + (uint32_t)(r0 + lane)) * P; };
                const unsigned char *const h1[1] = {pub};
                bool hit[1][1];
                freg_eval<NOFF, 1, 1>(fp, f, h1, real, myK, st, en, mb, x, my_hdr, par, hit, nk);
            }
            if (par[0] == (uint32_t)(nk + 1)) alive = false;   // covered by this lane (:120-122), or the lane itself
            ++nk;
            if ((long long)nk == sp.top_k) break;    // :133
        }

        // ---- 3. outputs -------------------------------------------------------------------------------------------------------------
        if (tid < N) {
            st_global_cs_u64(sp.parent + (size_t)f * N + tid, (long long)par[0]);
            if (tid >= nk) st_global_cs_u64(sp.keep + (size_t)f * N + tid, 0ll);   // :139-140
        }
        if (tid == 0) sp.num_keep[f] = sp.top_k < (long long)nk ? sp.top_k : (long long)nk;   // :142
        if (REC) {   // {keep[0 .. top_k), num} to every destination, one (column, destination) pair per thread
            const int w = rec.width;
            for (int i = tid; i < w * rec.n; i += (int)blockDim.x) {
                const int d = i / w, c = i - d * w;
                const long long v = c == w - 1 ? (sp.top_k < (long long)nk ? sp.top_k : (long long)nk) : (c < nk ? (long long)kept[c] : 0ll);
                rec.dst[d][(rec.row0 + f) * w + c] = v;
            }
        }
    }
}

}  // namespace phnms

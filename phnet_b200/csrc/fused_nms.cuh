// fused_nms.cuh -- the fast path: one thread-block CLUSTER per frame, proposals resident in shared memory.
//
// What it replaces (PHNet, paths relative to the reference repo):
//   libs/ops/csrc/nms.cpp:51              scores.sort(0, true)      -> rank keys + per-round cluster-wide arg-min (no sort)
//   libs/ops/csrc/nms_kernel.cu:26-48     devIoU                    -> pair_hit(): same fp32 op sequence, per pair sequential
//   libs/ops/csrc/nms_kernel.cu:50-96     nms_kernel (N^2 bitmask)  -> only the mask ROWS OF KEPT LANES are ever evaluated
//   libs/ops/csrc/nms_kernel.cu:99-143    nms_collect (<<<1,1>>>)   -> replicated greedy loop, one round per kept lane
//
// Why this is exact: nms_collect reads mask row i only when lane i is kept (:116-122), and bit j of row i is
// devIoU(lane i, lane j) for every j ranked after i (:85-91).  So for each kept lane, in rank order, we evaluate
// devIoU against every lower-ranked lane, mark hits as removed and stamp parent (last writer wins, :123-129).
// "removed or kept" == "parent != 0", so parent doubles as the removed set.
//
// Data movement per frame: the CTA's slab of proposal rows is pulled HBM -> shared memory once with TMA 1-D bulk
// copies (cp.async.bulk, UBLKCP) completing on an mbarrier; rows keep the reference's row-major [5+n_off] layout
// whose odd word stride (41 / 77) makes thread-per-row access bank-conflict free.  Outputs are written once,
// coalesced, with streaming stores.  Nothing else touches HBM.
#pragma once
#include "common.cuh"

namespace phnms {

struct FusedLayout {
    int off_wred;     // 32 x u64 per-warp arg-min scratch
    int off_bit;      // 32 x (float key, int val, int valid) bitonic scratch (n <= 32, torch sort model)
    int off_slots;    // 2 parities x csize candidate slots
    int slot_stride;  // bytes; slot = {key, gidx, start, end, row[round4(P)]}
    int off_colkey;   // rpc x u32
    int off_colse;    // rpc x int2 (start, end clamped)
    int off_colpar;   // rpc x u32 (0 = alive, else 1-based slot of the last covering kept lane)
    int off_rows;     // 16-byte aligned; 16 B lead + rpc*P*4 + 32 B tail/over-read pad
    int total;
};


struct FusedParams {
    const float *props;
    const float *scores;
    const int32_t *n_valid;
    long long *keep;
    long long *num_keep;
    long long *parent;
    long long F;
    long long top_k;
    int N;
    int n_off;
    int rpc;    // rows (proposals) per CTA
    int csize;  // CTAs per cluster == CTAs per frame
    int sort_model;
    float thr;
    FusedLayout L;
    const int *topm;   // optional [F, kTopM] candidate slots (header + row) of the best-ranked proposals (phnms_topm_kernel)
    int topm_count;    // candidate slots per frame in `topm` (<= kTopM)
    unsigned long long *claim_ctr;   // frame-claim counter (zero before the launch); register-resident kernel only
    const int *frame_list;           // optional: the kernel works through frame_list[0 .. *frame_count) instead of 0 .. F
    const unsigned int *frame_count; //           (the resume list of the streaming path, stream.cuh); static schedule only
    long long *trace;  // optional: CTA 0 / thread 0 writes clock64() at phase boundaries (phnms_forward_f32_trace)
    int trace_len;
    RecordSink rec;    // optional (rec.n > 0): the register-resident kernel also stores every frame's compact record
};

__host__ __device__ inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

inline FusedLayout fused_layout(int rpc, int P, int csize) {
    FusedLayout L;
    int o = 16;  // mbarrier at 0
    L.off_wred = o;
    o += 32 * 8;
    L.off_bit = o;
    o += 32 * 12;
    o = round_up(o, 16);
    L.off_slots = o;
    L.slot_stride = 16 + 4 * round_up(P, 4);
    o += 2 * csize * L.slot_stride;
    L.off_colkey = o;
    o += round_up(rpc * 4, 16);
    L.off_colse = o;
    o += round_up(rpc * 8, 16);
    L.off_colpar = o;
    o += round_up(rpc * 4, 16);
    L.off_rows = o;
    o += 16 + round_up(rpc * P * 4, 16) + 32;
    L.total = o;
    return L;
}

// One round of devIoU(kept lane a, every lower-ranked resident lane).
//   a      : kept lane's row in shared memory, 16-byte aligned, padded to a multiple of 4 words
//   wk     : kept lane's rank key (key << 32 | index); lanes with a larger key are ranked after it
__device__ __forceinline__ void fused_round(const FusedParams &p, const float *__restrict__ rows, const uint32_t *colkey,
                                            const int2 *colse, uint32_t *colpar, int r0, int nloc, const float *a,
                                            u64 wk, int sa, int ea, uint32_t slot1, int warp, int lane, int T) {
    const int P = 5 + p.n_off;
    for (int cb = warp * 32; cb < nloc; cb += T) {
        const int c = cb + lane;
        const bool valid = c < nloc;
        const uint32_t ck = valid ? colkey[c] : 0u;
        const u64 K = ((u64)ck << 32) | (uint32_t)(r0 + c);
        const bool act = valid && (K > wk);
        int2 se = make_int2(0, -1);
        if (act) se = colse[c];
        const float *b = rows + (size_t)(valid ? c : 0) * P;
        const bool hit = warp_pair_hit<true>(a, b, act, sa, ea, se.x, se.y, p.thr);
        if (hit || (valid && K == wk)) colpar[c] = slot1;  // nms_kernel.cu:127,:129
    }
}

__global__ void __launch_bounds__(512, 1) phnms_fused_kernel(const FusedParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    const int P = 5 + p.n_off;
    const int csize = p.csize;
    const uint32_t rank = csize > 1 ? cluster_ctarank() : 0u;
    const long long cl = blockIdx.x / csize, ncl = gridDim.x / csize;

    u64 *wred = reinterpret_cast<u64 *>(smem + p.L.off_wred);
    float *bit_key = reinterpret_cast<float *>(smem + p.L.off_bit);
    int *bit_val = reinterpret_cast<int *>(smem + p.L.off_bit + 128);
    int *bit_ok = reinterpret_cast<int *>(smem + p.L.off_bit + 256);
    unsigned char *slots = smem + p.L.off_slots;
    uint32_t *colkey = reinterpret_cast<uint32_t *>(smem + p.L.off_colkey);
    int2 *colse = reinterpret_cast<int2 *>(smem + p.L.off_colse);
    uint32_t *colpar = reinterpret_cast<uint32_t *>(smem + p.L.off_colpar);
    unsigned char *rows_buf = smem + p.L.off_rows;
    const uint32_t bar = smem_u32(smem);

    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (csize > 1) {  // every CTA of the cluster is resident before anyone writes into a peer's slots
        cluster_arrive_release();
        cluster_wait_acquire();
    }

    uint32_t load_phase = 0, round_ctr = 0;

    for (long long f = cl; f < p.F; f += ncl) {
        int nv = p.N;
        if (p.n_valid) nv = max(0, min(p.n_valid[f], p.N));
        const int r0 = min((int)rank * p.rpc, nv);
        const int nloc = min(p.rpc, nv - r0);

        // ---- HBM -> shared: this CTA's slab of rows --------------------------------------------------------
        const float *src = p.props + ((size_t)f * p.N + r0) * P;
        const uintptr_t s = (uintptr_t)src;
        const uintptr_t e = s + (size_t)nloc * P * 4;
        const uintptr_t s_al = (s + 15) & ~(uintptr_t)15, e_al = e & ~(uintptr_t)15;
        const bool bulk = e_al > s_al;
        const uint32_t head = bulk ? (uint32_t)(s_al - s) : 0u;  // 0,4,8,12 bytes in front of the aligned body
        float *rows = reinterpret_cast<float *>(rows_buf + 16 - head);
        if (bulk) {
            if (tid == 0) {
                const uint32_t total = (uint32_t)(e_al - s_al);
                mbar_arrive_expect_tx(bar, total);
                uint32_t chunk = ((total / 8 + 15) & ~15u);
                if (chunk < 4096u) chunk = 4096u;
                const uint32_t dst = smem_u32(rows_buf + 16);
                for (uint32_t off = 0; off < total; off += chunk)
                    bulk_g2s(dst + off, reinterpret_cast<const void *>(s_al + off), min(chunk, total - off), bar);
            }
            const int tail0 = (int)((e_al - s) >> 2), ntail = (int)((e - e_al) >> 2);
            if (tid < (int)(head >> 2)) rows[tid] = src[tid];
            if (tid >= 32 && tid - 32 < ntail) rows[tail0 + tid - 32] = src[tail0 + tid - 32];
        } else {
            for (int w = tid; w < nloc * P; w += T) rows[w] = src[w];
        }

        // ---- rank keys from scores (global loads overlap the bulk copy) ----------------------------------------
        const bool bitonic = (p.sort_model == 0) && nv <= 32 && nv >= 2;  // torch: unstable bitonic network (n <= 32)
        const bool nan_first = p.sort_model == 1;
        const float *sc = p.scores + (size_t)f * p.N + r0;
        for (int c = tid; c < nloc; c += T) {
            if (!bitonic) colkey[c] = key_desc(sc[c], nan_first);
            colpar[c] = 0u;
        }
        if (bitonic && warp == 0 && nloc > 0) {
            // ATen bitonicSortKVInPlace<block_dim_x = 16> (SortUtils.cuh), 32 slots, invalid slots sort last
            bit_ok[lane] = lane < nv;
            bit_key[lane] = lane < nv ? sc[lane] : 0.0f;
            bit_val[lane] = lane < nv ? lane : 0;
            __syncwarp();
            for (unsigned size = 2; size <= 32; size *= 2) {
                const bool last_merge = size == 32;
                const bool flag = !last_merge && ((lane & (size / 2)) != 0);
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
            if (lane < nv) colkey[bit_val[lane]] = (uint32_t)lane;  // rank position is the key
        }

        if (bulk) {
            mbar_wait(bar, load_phase);
            load_phase ^= 1u;
        }
        __syncthreads();

        // ---- per-lane [start, end] (nms_kernel.cu:29-34), once per proposal instead of once per pair ------------
        for (int c = tid; c < nloc; c += T) {
            const float *row = rows + (size_t)c * P;
            const int st = lane_start(row[2], p.n_off);
            colse[c] = make_int2(st, lane_end(row[4], st, p.n_off));
        }
        __syncthreads();

        // ---- greedy rounds: one per kept lane (nms_collect, :111-136) -------------------------------------------
        long long n = 0;
        while (true) {
            u64 best = kNone64;
            for (int c = tid; c < nloc; c += T)
                if (colpar[c] == 0u) best = min(best, ((u64)colkey[c] << 32) | (uint32_t)(r0 + c));
            best = warp_min_u64(best);
            if (lane == 0) wred[warp] = best;
            __syncthreads();
            best = warp_min_u64(lane < nwarps ? wred[lane] : kNone64);

            // publish this CTA's candidate {key, index, start, end, row} into slot[par][rank] of every CTA
            const uint32_t par = round_ctr & 1u;
            ++round_ctr;
            const bool have = best != kNone64;
            const int bc = have ? (int)((uint32_t)best - (uint32_t)r0) : 0;
            unsigned char *myslot = slots + (size_t)(par * csize + rank) * p.L.slot_stride;
            for (int w = tid; w < 4 + P; w += T) {
                uint32_t v;
                if (w == 0) v = (uint32_t)(best >> 32);
                else if (w == 1) v = (uint32_t)best;
                else if (w == 2) v = have ? (uint32_t)colse[bc].x : 0u;
                else if (w == 3) v = have ? (uint32_t)colse[bc].y : 0u;
                else v = have ? __float_as_uint(rows[(size_t)bc * P + (w - 4)]) : 0u;
                if (csize == 1) {
                    reinterpret_cast<uint32_t *>(myslot)[w] = v;
                } else {
                    const uint32_t addr = smem_u32(myslot) + 4u * w;
                    for (int d = 0; d < csize; ++d) st_cluster_u32(mapa_u32(addr, (uint32_t)d), v);
                }
            }
            if (csize > 1) {
                cluster_arrive_release();
                cluster_wait_acquire();
            } else {
                __syncthreads();
            }

            // the winner over the cluster: smallest (key, index) == first not-removed lane in sorted order (:116)
            u64 wk = kNone64;
            int wslot = 0;
            for (int d = 0; d < csize; ++d) {
                const uint2 h = *reinterpret_cast<const uint2 *>(slots + (size_t)(par * csize + d) * p.L.slot_stride);
                const u64 k = ((u64)h.x << 32) | h.y;
                if (k < wk) { wk = k; wslot = d; }
            }
            if (wk == kNone64) break;  // every lane is kept or removed
            const unsigned char *ws = slots + (size_t)(par * csize + wslot) * p.L.slot_stride;
            const int2 sea = *reinterpret_cast<const int2 *>(ws + 8);
            if (rank == 0 && tid == 0) p.keep[(size_t)f * p.N + n] = (long long)(uint32_t)wk;  // :118

            fused_round(p, rows, colkey, colse, colpar, r0, nloc, reinterpret_cast<const float *>(ws + 16), wk, sea.x,
                        sea.y, (uint32_t)(n + 1), warp, lane, T);
            ++n;
            if (n == p.top_k) break;  // :133 (top_k == 0 never stops early)
            __syncthreads();
        }
        __syncthreads();

        // ---- outputs, written once: parent, zero padding of keep, count ------------------------------------------
        {
            const int o0 = (int)rank * p.rpc, o1 = min(o0 + p.rpc, p.N);
            long long *keep_f = p.keep + (size_t)f * p.N, *par_f = p.parent + (size_t)f * p.N;
            for (int i = o0 + tid; i < o1; i += T) {
                st_global_cs_u64(par_f + i, i < nv ? (long long)colpar[i - o0] : 0ll);
                if (i >= n) st_global_cs_u64(keep_f + i, 0ll);  // :139-140
            }
            if (rank == 0 && tid == 0) p.num_keep[f] = p.top_k < n ? p.top_k : n;  // :142
        }
        __syncthreads();  // rows / col* are rewritten by the next frame
    }

    if (csize > 1) {  // no CTA leaves while a peer may still address its shared memory
        cluster_arrive_release();
        cluster_wait_acquire();
    }
}

}  // namespace phnms

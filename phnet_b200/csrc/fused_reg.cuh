// fused_reg.cuh -- the fast path for the two offset counts PHNet ships (36 and 72): proposals live in REGISTERS.
//
// Same algorithm and exactness argument as fused_nms.cuh (one cluster of CTAs per frame, one greedy round per kept
// lane, only the mask rows nms_collect would read are evaluated), restructured around what the ncu profile of the
// shared-memory version showed (profiles/r1_v1_*):
//
//   * each thread owns CPT proposals and holds their full rows (5+NOFF floats) in registers.  A row is read from
//     shared memory exactly once per frame instead of once per round, and the offset loop is fully unrolled:
//     per offset  FADD (a-x)  +  LOP3 (in-range bit -> predicate)  +  predicated FADD (dist += |.|)   ~3.25 instr
//     instead of 7.5 (two compares, index arithmetic, select, one LDS).
//   * shared memory is then only a STAGING buffer for the TMA bulk copy.  It is free as soon as the rows are in
//     registers, so the next frame's slab is requested immediately: HBM traffic overlaps the greedy rounds.
//   * the per-round exchange between the CTAs of a cluster uses st.async (remote shared-memory stores that complete
//     a transaction count on the RECEIVER's mbarrier) instead of barrier.cluster, whose release/acquire compiles to
//     MEMBAR.ALL.GPU + ERRBAR + UCGABAR and was ~20 % of all stall samples.
//
// Reference semantics: libs/ops/csrc/nms.cpp:51 (ordering), nms_kernel.cu:26-48 (devIoU), :50-96 (mask), :99-143 (collect).
#pragma once
#include "common.cuh"
#include "fused_nms.cuh"

namespace phnms {

struct FregLayout {
    int off_wred;     // 2 x 32 x u64 (double buffered per round parity)
    int off_bit;      // bitonic scratch
    int off_slots;    // 2 parities x csize slots
    int slot_stride;  // 16 + 4 * round4(P)
    int off_rows;     // staging: 16 B lead + rpc*P*4 + pad
    int total;
};

inline FregLayout freg_layout(int rpc, int P, int csize) {
    FregLayout L;
    int o = 32;  // mbarriers: load @0, exchange @8 and @16
    L.off_wred = o;
    o += 2 * 32 * 8;
    L.off_bit = o;
    o += 32 * 12;
    o = round_up(o, 16);
    L.off_slots = o;
    L.slot_stride = 16 + 4 * round_up(P, 4);
    o += 2 * csize * L.slot_stride;
    L.off_rows = o;
    o += 16 + round_up(rpc * P * 4, 16) + 32;
    L.total = o;
    return L;
}

__device__ __forceinline__ void st_async_v4(uint32_t dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(dst),
                 "r"(a), "r"(b), "r"(c), "r"(d), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void sts_v4(uint32_t dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

struct Slab {       // what one CTA loads for one frame
    int nv;         // real proposals in the frame
    int r0;         // first row owned by this CTA
    int nloc;       // rows owned
    int head;       // bytes between the row data and the 16-byte aligned bulk body (0,4,8,12)
    bool bulk;      // a TMA bulk copy is in flight for it
};

// Requests frame f's slab: TMA bulk copy of the 16-byte aligned body (completes on `bar`), the <= 3 unaligned words at
// either end by ordinary loads.  Call with the staging buffer free (after a __syncthreads that follows its last read).
__device__ __forceinline__ Slab request_slab(const FusedParams &p, long long f, uint32_t rank, unsigned char *rows_buf,
                                             uint32_t bar, int tid, int T, int P) {
    Slab s;
    s.nv = p.N;
    if (p.n_valid) s.nv = max(0, min(p.n_valid[f], p.N));
    s.r0 = min((int)rank * p.rpc, s.nv);
    s.nloc = min(p.rpc, s.nv - s.r0);
    const float *src = p.props + ((size_t)f * p.N + s.r0) * P;
    const uintptr_t b = (uintptr_t)src, e = b + (size_t)s.nloc * P * 4;
    const uintptr_t b_al = (b + 15) & ~(uintptr_t)15, e_al = e & ~(uintptr_t)15;
    s.bulk = e_al > b_al;
    s.head = s.bulk ? (int)(b_al - b) : 0;
    float *rows = reinterpret_cast<float *>(rows_buf + 16 - s.head);
    if (s.bulk) {
        if (tid == 0) {
            const uint32_t total = (uint32_t)(e_al - b_al);
            fence_proxy_async();
            mbar_arrive_expect_tx(bar, total);
            uint32_t chunk = ((total / 8 + 15) & ~15u);
            if (chunk < 4096u) chunk = 4096u;
            const uint32_t dst = smem_u32(rows_buf + 16);
            for (uint32_t off = 0; off < total; off += chunk)
                bulk_g2s(dst + off, reinterpret_cast<const void *>(b_al + off), min(chunk, total - off), bar);
        }
        const int tail0 = (int)((e_al - b) >> 2), ntail = (int)((e - e_al) >> 2);
        if (tid >= 32 && tid - 32 < (s.head >> 2)) rows[tid - 32] = src[tid - 32];
        if (tid >= 64 && tid - 64 < ntail) rows[tail0 + tid - 64] = src[tail0 + tid - 64];
    } else {
        for (int w = tid; w < s.nloc * P; w += T) rows[w] = src[w];
    }
    return s;
}

template <int NOFF, int CPT>
__global__ void __launch_bounds__(512, 1) phnms_freg_kernel(const FusedParams p, const FregLayout L) {
    constexpr int P = 5 + NOFF;
    constexpr int MW = (P + 31) / 32;   // in-range bitmask words
    constexpr int P4 = (P + 3) & ~3;
    constexpr int SLOT = 16 + 4 * P4;   // bytes: {key, index, start, end} + padded row

    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    const int csize = p.csize;
    const uint32_t rank = csize > 1 ? cluster_ctarank() : 0u;
    const long long cl = blockIdx.x / csize, ncl = gridDim.x / csize;

    u64 *wred = reinterpret_cast<u64 *>(smem + L.off_wred);
    float *bit_key = reinterpret_cast<float *>(smem + L.off_bit);
    int *bit_val = reinterpret_cast<int *>(smem + L.off_bit + 128);
    int *bit_ok = reinterpret_cast<int *>(smem + L.off_bit + 256);
    unsigned char *slots = smem + L.off_slots;
    unsigned char *rows_buf = smem + L.off_rows;
    const uint32_t bar_load = smem_u32(smem), bar_x0 = bar_load + 8;

    if (tid == 0) {
        mbar_init(bar_load, 1);
        mbar_init(bar_x0, 1);
        mbar_init(bar_x0 + 8, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (csize > 1) {  // all CTAs of the cluster are resident and their mbarriers initialised before any remote store
        cluster_arrive_release();
        cluster_wait_acquire();
    }

    uint32_t load_phase = 0, round_ctr = 0;
    float sc[CPT];       // this frame's scores of my columns
    Slab cur;
    if (cl < p.F) {
        cur = request_slab(p, cl, rank, rows_buf, bar_load, tid, T, P);
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const int col = c * T + tid;
            sc[c] = col < cur.nloc ? p.scores[(size_t)cl * p.N + cur.r0 + col] : 0.0f;
        }
    }

    for (long long f = cl; f < p.F; f += ncl) {
        // ---- staging -> registers -------------------------------------------------------------------------------
        if (cur.bulk) {
            mbar_wait(bar_load, load_phase);
            load_phase ^= 1u;
        }
        __syncthreads();
        const float *rows = reinterpret_cast<const float *>(rows_buf + 16 - cur.head);
        const int nv = cur.nv, r0 = cur.r0, nloc = cur.nloc;
        float x[CPT][P];
        bool valid[CPT];
        uint32_t key[CPT], par[CPT];
        int st[CPT], en[CPT];
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const int col = c * T + tid;
            valid[c] = col < nloc;
            const float *row = rows + (size_t)(valid[c] ? col : 0) * P;
#pragma unroll
            for (int i = 0; i < P; ++i) x[c][i] = row[i];
            st[c] = lane_start(x[c][2], NOFF);           // nms_kernel.cu:29-30
            en[c] = lane_end(x[c][4], st[c], NOFF);      // :32-34
            key[c] = key_desc(sc[c], p.sort_model == 1);
            par[c] = 0u;
        }
        const bool bitonic = (p.sort_model == 0) && nv <= 32 && nv >= 2;  // torch: unstable bitonic network (n <= 32)
        if (bitonic && warp == 0 && nloc > 0) {  // rank 0 holds the whole frame (rows_per_cta >= 32), column == lane
            bit_ok[lane] = lane < nv;
            bit_key[lane] = lane < nv ? sc[0] : 0.0f;
            bit_val[lane] = lane < nv ? lane : 0;
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
            // sorted position -> rank key of the proposal that landed there
            int mypos = 0;
            for (int q = 0; q < 32; ++q)
                if (bit_val[q] == lane && q < nv) mypos = q;
            key[0] = (uint32_t)mypos;
        }
        __syncthreads();  // every row is in registers: the staging buffer is free

        // ---- request the next frame now; it lands while this frame's rounds run ----------------------------------
        Slab nxt;
        nxt.bulk = false; nxt.nv = nxt.r0 = nxt.nloc = nxt.head = 0;
        const long long fn = f + ncl;
        if (fn < p.F) {
            nxt = request_slab(p, fn, rank, rows_buf, bar_load, tid, T, P);
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const int col = c * T + tid;
                sc[c] = col < nxt.nloc ? p.scores[(size_t)fn * p.N + nxt.r0 + col] : 0.0f;
            }
        }

        u64 myK[CPT];
#pragma unroll
        for (int c = 0; c < CPT; ++c)
            myK[c] = valid[c] ? (((u64)key[c] << 32) | (uint32_t)(r0 + c * T + tid)) : kNone64;

        // ---- greedy rounds: one per kept lane (nms_collect, :111-136) -------------------------------------------
        long long n = 0;
        while (true) {
            const uint32_t par_bit = round_ctr & 1u, xphase = (round_ctr >> 1) & 1u;
            ++round_ctr;
            u64 best = kNone64;
#pragma unroll
            for (int c = 0; c < CPT; ++c)
                if (par[c] == 0u) best = min(best, myK[c]);
            best = warp_min_u64(best);
            if (lane == 0) wred[par_bit * 32 + warp] = best;
            __syncthreads();
            best = warp_min_u64(lane < nwarps ? wred[par_bit * 32 + lane] : kNone64);

            // publish the CTA's candidate {key, index, start, end, row} into slot[par][rank] of every CTA
            const uint32_t myslot = smem_u32(slots + (size_t)(par_bit * csize + rank) * L.slot_stride);
            const uint32_t bar_x = bar_x0 + 8u * par_bit;
            if (csize > 1 && tid == 0) mbar_arrive_expect_tx(bar_x, (uint32_t)(csize * SLOT));
            bool owner = false;
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                if (myK[c] == best && best != kNone64) {
                    owner = true;
                    if (csize == 1) {
                        sts_v4(myslot, (uint32_t)(best >> 32), (uint32_t)best, (uint32_t)st[c], (uint32_t)en[c]);
#pragma unroll
                        for (int g = 0; g < P4 / 4; ++g)
                            sts_v4(myslot + 16 + 16 * g, __float_as_uint(x[c][4 * g]),
                                   4 * g + 1 < P ? __float_as_uint(x[c][4 * g + 1 < P ? 4 * g + 1 : 0]) : 0u,
                                   4 * g + 2 < P ? __float_as_uint(x[c][4 * g + 2 < P ? 4 * g + 2 : 0]) : 0u,
                                   4 * g + 3 < P ? __float_as_uint(x[c][4 * g + 3 < P ? 4 * g + 3 : 0]) : 0u);
                    } else {
                        for (int d = 0; d < csize; ++d) {
                            const uint32_t dst = mapa_u32(myslot, (uint32_t)d), dbar = mapa_u32(bar_x, (uint32_t)d);
                            st_async_v4(dst, (uint32_t)(best >> 32), (uint32_t)best, (uint32_t)st[c], (uint32_t)en[c], dbar);
#pragma unroll
                            for (int g = 0; g < P4 / 4; ++g)
                                st_async_v4(dst + 16 + 16 * g, __float_as_uint(x[c][4 * g]),
                                            4 * g + 1 < P ? __float_as_uint(x[c][4 * g + 1 < P ? 4 * g + 1 : 0]) : 0u,
                                            4 * g + 2 < P ? __float_as_uint(x[c][4 * g + 2 < P ? 4 * g + 2 : 0]) : 0u,
                                            4 * g + 3 < P ? __float_as_uint(x[c][4 * g + 3 < P ? 4 * g + 3 : 0]) : 0u, dbar);
                        }
                    }
                }
            }
            if (best == kNone64 && tid == 0) {  // nothing alive here: an empty slot of the same size
                if (csize == 1) {
                    sts_v4(myslot, 0xffffffffu, 0xffffffffu, 0u, 0u);
                } else {
                    for (int d = 0; d < csize; ++d) {
                        const uint32_t dst = mapa_u32(myslot, (uint32_t)d), dbar = mapa_u32(bar_x, (uint32_t)d);
                        st_async_v4(dst, 0xffffffffu, 0xffffffffu, 0u, 0u, dbar);
                        for (int g = 0; g < P4 / 4; ++g) st_async_v4(dst + 16 + 16 * g, 0u, 0u, 0u, 0u, dbar);
                    }
                }
            }
            (void)owner;
            if (csize > 1) mbar_wait(bar_x, xphase);
            else __syncthreads();

            // the winner over the cluster: smallest (key, index) == first not-removed lane in sorted order (:116)
            u64 wk = kNone64;
            int wslot = 0;
            for (int d = 0; d < csize; ++d) {
                const uint2 h = *reinterpret_cast<const uint2 *>(slots + (size_t)(par_bit * csize + d) * L.slot_stride);
                const u64 k = ((u64)h.x << 32) | h.y;
                if (k < wk) { wk = k; wslot = d; }
            }
            if (wk == kNone64) break;  // every lane is kept or removed
            const unsigned char *ws = slots + (size_t)(par_bit * csize + wslot) * L.slot_stride;
            const int2 sea = *reinterpret_cast<const int2 *>(ws + 8);
            const uint32_t a_addr = smem_u32(ws + 16);
            if (rank == 0 && tid == 0) p.keep[(size_t)f * p.N + n] = (long long)(uint32_t)wk;  // :118

            // devIoU(kept lane, my lanes): in-range bitmask per lane, fully unrolled ascending sum (:38-44)
            uint32_t m[CPT][MW];
            float dist[CPT];
            bool act[CPT];
            int len[CPT];
            bool any_act = false;
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const int start = max(sea.x, st[c]);   // :31
                const int end = min(sea.y, en[c]);     // :34 (both clamped to NOFF-1)
                act[c] = valid[c] && (myK[c] > wk) && (end >= start);  // :36
                const int i0 = (int)(((uint32_t)start + 5u) & 255u);   // :38 unsigned char counter
                const int last = (int)((uint32_t)end + 5u);
                const bool run = act[c] && (i0 <= last);
                len[c] = (int)((uint32_t)end - (uint32_t)start + 1u);
#pragma unroll
                for (int w = 0; w < MW; ++w) {
                    const int l = max(i0 - 32 * w, 0), h = min(last - 32 * w, 31);
                    m[c][w] = (run && l <= h) ? ((0xffffffffu >> (31 - h)) & (0xffffffffu << l)) : 0u;
                }
                dist[c] = 0.0f;
                any_act |= run;
            }
            if (__any_sync(0xffffffffu, any_act)) {
#pragma unroll
                for (int g = 0; g < P4 / 4; ++g) {
                    const float4 av = lds_v4(a_addr + 16 * g);
                    const float a4[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = 4 * g + u;
                        if (i < P) {
#pragma unroll
                            for (int c = 0; c < CPT; ++c) {
                                const float t = __fsub_rn(a4[u], x[c][i]);
                                if (m[c][i >> 5] & (1u << (i & 31))) dist[c] = __fadd_rn(dist[c], fabsf(t));
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const bool hit = act[c] && (dist[c] < __fmul_rn(p.thr, (float)len[c]));  // :46
                if (hit || myK[c] == wk) par[c] = (uint32_t)(n + 1);                      // :127,:129
            }
            ++n;
            if (n == p.top_k) break;  // :133 (top_k == 0 never stops early)
        }

        // ---- outputs, written once: parent, zero padding of keep, count ------------------------------------------
        {
            const int o0 = (int)rank * p.rpc, o1 = min(o0 + p.rpc, p.N);
            long long *keep_f = p.keep + (size_t)f * p.N, *par_f = p.parent + (size_t)f * p.N;
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const int i = o0 + c * T + tid;
                if (i < o1) {
                    st_global_cs_u64(par_f + i, (long long)par[c]);
                    if (i >= n) st_global_cs_u64(keep_f + i, 0ll);  // :139-140
                }
            }
            if (rank == 0 && tid == 0) p.num_keep[f] = p.top_k < n ? p.top_k : n;  // :142
        }
        cur = nxt;
    }

    if (csize > 1) {  // no CTA leaves while a peer may still address its shared memory
        cluster_arrive_release();
        cluster_wait_acquire();
    }
}

}  // namespace phnms

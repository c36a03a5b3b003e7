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
//   * the first batch of a frame is PLANNED by phnms_topm_kernel (candidates, their mutual predicate, the greedy scan
//     over them, their keep[] entries): the kept candidates are known up front and evaluated two per pass over the
//     registers (freg_eval_multi: independent fp32 chains, FADD2 subtractions) with no selection, exchange or barrier.
//     Only frames whose candidates run out before top_k lanes are kept enter the fallback batches (cluster exchange).
//   * what bounds it now: latency inside the CTA -- the phases of a frame run one after the other in
//     all 16 warps, 4 warps per scheduler because 72 offsets live in registers (128 registers / thread, zero spills:
//     every additional live value in the frame loop spills and costs 10-15 %).
//
// Reference semantics: libs/ops/csrc/nms.cpp:51 (ordering), nms_kernel.cu:26-48 (devIoU), :50-96 (mask), :99-143 (collect).
#pragma once
#include "common.cuh"
#include "fused_nms.cuh"

namespace phnms {

constexpr int kCand = 4;   // candidates every CTA publishes per exchange (fallback batches)
constexpr int kTopM = 16;  // capacity: per-frame best-ranked proposals precomputed by phnms_topm_kernel (first batch);
                           // FusedParams::topm_count (8 or 12 for top_k <= 4, else 16) of them are produced and fetched
constexpr int kPlanLanes = 2;    // kept lanes of the planned batch evaluated per pass, one proposal per thread (n_off 72 / small frames)
constexpr int kPlanLanes2 = 2;   // same, two proposals per thread (n_off 36): chains per thread = lanes x 2
constexpr bool kSkipGroups = true;   // warp-uniform skip of 8-word groups outside the kept lane's own range
constexpr bool kPackedSub = true;    // a - x of two neighbouring offsets as one FADD2 (sub.f32x2)
constexpr int kHdr = 32;   // candidate header bytes: {key, index, start, end, mask0, mask1, mask2, aux}

// A candidate slot = 32-byte header + the proposal's row padded to a multiple of 4 words.
//   key/index : rank key (key << 32 | index orders the frame), original proposal index
//   start/end : the lane's own bounds (nms_kernel.cu:29-34), end clamped to n_off-1
//   mask0..2  : the lane's in-range bitmask over row words (range_mask)
//   aux       : candidate block: adjacency bits << 16 | number of valid candidates (slot 1: | the kept set of the planned
//               greedy scan instead of the count); compacted fallback headers: index of the slot that holds the row
struct FregLayout {
    int off_wtop;     // 2 parities x 32 warps x kCand x u64: per-warp best alive keys
    int off_bh;       // 32 x 32 B compacted headers of a fallback batch, in rank order; then count; then 3 x 32 dead flags
    int off_bit;      // bitonic scratch
    int off_slots;    // 2 parities x csize x kCand exchange slots
    int slot_stride;  // kHdr + 4 * round4(P)
    int off_pslots;   // 2 frame parities x kTopM slots: precomputed candidates of the current / next frame
    int off_sbuf;     // 2 frame parities x rpc scores (this CTA's slice), filled by cp.async one frame ahead
    int off_rows;     // staging: 16 B lead + rpc*P*4 + pad
    int total;
};

inline FregLayout freg_layout(int rpc, int P, int csize) {
    FregLayout L;
    int o = 96;  // mbarriers: load @0, exchange @8 / @16, frame claim @24 / @32; claimed frame indices @40 / @48;
                 // slab geometry record (Slab, 32 bytes) @64
    L.off_wtop = o;
    o += 2 * 32 * kCand * 8;
    L.off_bh = o;
    o += 32 * kHdr + 16 + 3 * 32 * 4;
    L.off_bit = o;
    o += 32 * 12;
    o = round_up(o, 16);
    L.off_slots = o;
    L.slot_stride = kHdr + 4 * round_up(P, 4);
    o += 2 * csize * kCand * L.slot_stride;
    L.off_pslots = o;
    o += 2 * kTopM * L.slot_stride;
    L.off_sbuf = o;
    o += 2 * round_up(rpc * 4, 16);
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

// {d0, d1} = {a0 - b0, a1 - b1}: one packed fp32x2 instruction (FADD2), each half rounded to nearest like FADD
__device__ __forceinline__ void fsub2(float a0, float a1, float b0, float b1, float &d0, float &d1) {
    asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tsub.f32x2 rd, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d0), "=f"(d1)
        : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}

__device__ __forceinline__ void cp_async_4(uint32_t dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_plain() { asm volatile("barrier.cluster.wait.aligned;" ::: "memory"); }
__device__ __forceinline__ void st_async_b64(uint32_t dst, unsigned long long v, uint32_t bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(dst), "l"(v), "r"(bar)
                 : "memory");
}

// Bits [5+start, 5+end] of one lane's own offset range (start below -5 acts as "no lower bound": the other lane's
// start wins the max).  Bit i corresponds to row word i (columns 0..4 are the header, nms_kernel.cu:38).
template <int MW>
__device__ __forceinline__ void range_mask(int start, int end, uint32_t (&m)[MW]) {
    const int lo = start < -5 ? 0 : min(start, 1 << 20) + 5;   // (clamps keep the int arithmetic below from wrapping)
    const int hi = max(end, -(1 << 20)) + 5;                   // end is already clamped to NOFF-1 from above
#pragma unroll
    for (int w = 0; w < MW; ++w) {
        const int l = max(lo - 32 * w, 0), h = min(hi - 32 * w, 31);
        m[w] = (l <= h) ? ((0xffffffffu >> (31 - h)) & (0xffffffffu << l)) : 0u;
    }
}

// What one CTA loads for one frame: its slab of rows, cut into a 16-byte aligned body (TMA bulk copies) and <= 3 words at
// either end (plain loads).  A pure function of the frame index.  When every frame starts at the same address modulo 16
// and has the same number of proposals (n_valid == nullptr and N * P * 4 a multiple of 16: the batch tensors PHNet and the
// bench hand over), the record is the same for every frame: thread 0 computes it once, everybody reads it back from
// shared memory (two LDS.128) instead of redoing ~100 instructions of 64-bit address arithmetic per warp and frame.
struct Slab {
    int nv;          // real proposals in the frame
    int r0;          // first row owned by this CTA
    int nloc;        // rows owned
    int head;        // bytes between the row data and the 16-byte aligned bulk body (0,4,8,12)
    uint32_t total;  // bytes of the aligned body (0: no bulk copy for the rows)
    uint32_t chunk;  // the body is requested in pieces of this size, one per issuing warp
    int tail0;       // first word (relative to the slab) after the aligned body
    int ntail;       // words after the aligned body
};

__device__ __forceinline__ int slab_issuers(int T) { return (T >> 5) >= 8 ? 8 : 4; }

__device__ __forceinline__ Slab slab_compute(const FusedParams &p, long long f, uint32_t rank, int P, int T) {
    Slab s;
    s.nv = p.N;
    if (p.n_valid) s.nv = max(0, min(p.n_valid[f], p.N));
    s.r0 = min((int)rank * p.rpc, s.nv);
    s.nloc = min(p.rpc, s.nv - s.r0);
    const uintptr_t b = (uintptr_t)(p.props + ((size_t)f * p.N + s.r0) * P), e = b + (size_t)s.nloc * P * 4;
    const uintptr_t b_al = (b + 15) & ~(uintptr_t)15, e_al = e & ~(uintptr_t)15;
    const bool body = e_al > b_al;
    s.head = body ? (int)(b_al - b) : 0;
    s.total = body ? (uint32_t)(e_al - b_al) : 0u;
    uint32_t chunk = ((s.total / (uint32_t)slab_issuers(T) + 15u) & ~15u);
    s.chunk = chunk < 2048u ? 2048u : chunk;
    s.tail0 = (int)((e_al - b) >> 2);
    s.ntail = (int)((e - e_al) >> 2);
    return s;
}

__device__ __forceinline__ bool slab_is_uniform(const FusedParams &p, int P) {
    return p.n_valid == nullptr && (((size_t)p.N * P * 4) & 15u) == 0u;
}

__device__ __forceinline__ Slab slab_geometry(const FusedParams &p, long long f, uint32_t rank, int P, int T, bool uniform,
                                              const unsigned char *geo) {
    if (!uniform) return slab_compute(p, f, rank, P, T);
    const int4 g0 = *reinterpret_cast<const int4 *>(geo), g1 = *reinterpret_cast<const int4 *>(geo + 16);
    Slab s;
    s.nv = g0.x; s.r0 = g0.y; s.nloc = g0.z; s.head = g0.w;
    s.total = (uint32_t)g1.x; s.chunk = (uint32_t)g1.y; s.tail0 = g1.z; s.ntail = g1.w;
    return s;
}

// Requests frame f's slab: TMA bulk copy of the 16-byte aligned body (completes on `bar`), the <= 3 unaligned words at
// either end by ordinary loads.  Call with the staging buffer free (after a __syncthreads that follows its last read).
// When phnms_topm_kernel ran, the frame's candidate block (kTopM slots, laid out exactly like pslots) rides along as
// one more bulk copy on the same mbarrier.
__device__ __forceinline__ void request_slab(const FusedParams &p, long long f, const Slab &s, unsigned char *rows_buf,
                                             uint32_t bar, int tid, int T, int P, unsigned char *cand_dst,
                                             uint32_t cand_bytes, float *sbuf, bool sbulk) {
    const float *src = p.props + ((size_t)f * p.N + s.r0) * P;
    const bool body = s.total != 0u;
    const bool cand = p.topm != nullptr;
    float *rows = reinterpret_cast<float *>(rows_buf + 16 - s.head);
    // this CTA's slice of the scores: one more bulk copy when it is 16-byte aligned (`sbulk`), else asynchronous 4-byte
    // copies (no register is tied up while they are in flight)
    const uint32_t sbytes = sbulk ? (((uint32_t)s.nloc * 4u + 15u) & ~15u) : 0u;
    if (sbulk) {
        if (tid == 64 && sbytes) bulk_g2s(smem_u32(sbuf), p.scores + (size_t)f * p.N + s.r0, sbytes, bar);
    } else {
        for (int c = tid; c < s.nloc; c += T) cp_async_4(smem_u32(sbuf + c), p.scores + (size_t)f * p.N + s.r0 + c);
        cp_async_commit();
    }
    if (cand) {
        if (tid == 0) mbar_arrive_expect_tx(bar, s.total + cand_bytes + sbytes);
        if (tid == 32)
            bulk_g2s(smem_u32(cand_dst), reinterpret_cast<const unsigned char *>(p.topm) + (size_t)f * cand_bytes, cand_bytes, bar);
    } else if (sbytes && !body) {
        if (tid == 0) mbar_arrive_expect_tx(bar, sbytes);
    }
    if (body) {
        // No proxy fence: the generic-proxy READS of the staging buffer are ordered before this async-proxy write by
        // the __syncthreads that precedes the call (a fence here compiles to MEMBAR and waits for the previous frame's
        // streaming stores to HBM: 2.3k cycles per frame in the phase trace).  The copy is cut into up to 8 chunks,
        // each issued by lane 0 of a different warp (back-to-back UBLKCPs from one thread serialise on the TMA queue
        // and sat on the critical path); a chunk may complete before thread 0 has armed the barrier, which is fine.
        if (tid == 0 && !cand) mbar_arrive_expect_tx(bar, s.total + sbytes);
        {
            const int nw = T >> 5, issuers = slab_issuers(T);
            const int w = (tid >> 5), k = nw - 1 - w;   // last warps issue: warp 0 has the first-batch work ahead of it
            if ((tid & 31) == 0 && k < issuers) {
                const uint32_t off = (uint32_t)k * s.chunk;
                if (off < s.total)
                    bulk_g2s(smem_u32(rows_buf + 16) + off, reinterpret_cast<const unsigned char *>(src) + s.head + off,
                             min(s.chunk, s.total - off), bar);
            }
        }
        if (tid >= 32 && tid - 32 < (s.head >> 2)) rows[tid - 32] = src[tid - 32];
        if (tid >= 64 && tid - 64 < s.ntail) rows[s.tail0 + tid - 64] = src[s.tail0 + tid - 64];
    } else {
        for (int w = tid; w < s.nloc * P; w += T) rows[w] = src[w];
    }
}

// devIoU of NK kept lanes (headers hdr[k] in shared memory, rows 32 bytes behind them) against this thread's lanes, in
// ONE pass over the registers: NK independent fp32 chains, each the reference's ascending sequential sum (:38-44).
// A lane's in-range bitmask for a pair is the AND of the two per-lane masks ([max(sa,sb), min(ea,eb)] is the
// intersection of the two ranges) -- unless BOTH starts are below -5, where the reference's unsigned-char counter wraps
// (rare; recomputed from the wrapped counter).  A start in [-5,-1] pulls header words 0..4 into the sum; they are not in
// registers and are re-read (rare).  parent is stamped in keep order, so the last kept lane that covers a lane wins.
template <int NOFF, int CPT, int NK, typename HdrFn>
__device__ __forceinline__ void freg_eval(const FusedParams &p, long long f, const unsigned char *const (&hdr)[NK],
                                          const bool (&live)[CPT], const u64 (&myK)[CPT], const int (&st)[CPT],
                                          const int (&en)[CPT], const uint32_t (&mb)[CPT][(5 + NOFF + 31) / 32],
                                          const float (&x)[CPT][NOFF], HdrFn my_hdr,
                                          uint32_t (&par)[CPT], bool (&hit_out)[NK][CPT], long long n_base) {
    constexpr int P = 5 + NOFF, MW = (P + 31) / 32, P4 = (P + 3) & ~3;
    u64 wk[NK];
    uint32_t a_addr[NK];
    uint32_t m[NK][CPT][MW];
    float dist[NK][CPT];
    bool act[NK][CPT];
    int len[NK][CPT];
    bool any_act = false, hdr_terms = false;
    uint32_t um_l[MW];   // union of the kept lanes' own in-range masks (the same value in every thread)
#pragma unroll
    for (int w = 0; w < MW; ++w) um_l[w] = 0u;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
        const uint4 wh = *reinterpret_cast<const uint4 *>(hdr[k]);
        const uint4 wm = *reinterpret_cast<const uint4 *>(hdr[k] + 16);
        wk[k] = ((u64)wh.x << 32) | wh.y;
        const int sa = (int)wh.z, ea = (int)wh.w;
        const uint32_t ma[3] = {wm.x, wm.y, wm.z};
        a_addr[k] = smem_u32(hdr[k] + kHdr);
#pragma unroll
        for (int w = 0; w < MW; ++w) um_l[w] |= ma[w];
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const int start = max(sa, st[c]);   // nms_kernel.cu:31
            const int end = min(ea, en[c]);     // :34 (both clamped to NOFF-1)
            act[k][c] = live[c] && (myK[c] > wk[k]) && (end >= start);  // :36
            len[k][c] = (int)((uint32_t)end - (uint32_t)start + 1u);
#pragma unroll
            for (int w = 0; w < MW; ++w) m[k][c][w] = act[k][c] ? (ma[w] & mb[c][w]) : 0u;
            if (act[k][c] && start < -5) {      // :38 unsigned char counter wrapped: (5 + start) & 255
                const int i0 = (int)(((uint32_t)start + 5u) & 255u), last = (int)((uint32_t)end + 5u);
                const bool run = i0 <= last;    // then 0 <= last <= NOFF+4: nothing below can wrap
#pragma unroll
                for (int w = 0; w < MW; ++w) {
                    const int l = max(i0 - 32 * w, 0), hh = min((run ? last : -1) - 32 * w, 31);
                    m[k][c][w] = (run && l <= hh) ? ((0xffffffffu >> (31 - hh)) & (0xffffffffu << l)) : 0u;
                }
            }
            dist[k][c] = 0.0f;
            any_act |= act[k][c];
            hdr_terms |= (m[k][c][0] & 0x1fu) != 0u;
        }
    }
    if (__any_sync(0xffffffffu, any_act)) {
        uint32_t um[MW];
#pragma unroll
        for (int w = 0; w < MW; ++w) um[w] = kSkipGroups ? __reduce_or_sync(0xffffffffu, um_l[w]) : 0xffffffffu;
        if (__any_sync(0xffffffffu, hdr_terms)) {
#pragma unroll
            for (int k = 0; k < NK; ++k) {
                const float *arow = reinterpret_cast<const float *>(hdr[k] + kHdr);
#pragma unroll
                for (int c = 0; c < CPT; ++c)
                    if (m[k][c][0] & 0x1fu) {
                        const float *mine = my_hdr(c);
                        for (int i = 0; i < 5; ++i)
                            if (m[k][c][0] & (1u << i))
                                dist[k][c] = __fadd_rn(dist[k][c], fabsf(__fsub_rn(arow[i], mine[i])));
                    }
            }
        }
        // Row words are walked in groups of 8 (two LDS.128 of the kept row).  A group in which no kept lane of this pass
        // has an in-range word is skipped by the whole warp: the pair mask is a subset of the kept lane's own mask
        // (also on the wrapped-counter path: [i0, last] lies inside [0, 5+ea]), so nothing would be added there.  `um`
        // comes out of REDUX in a uniform register, which makes these warp-uniform branches (no reconvergence code).
        // The subtractions of two neighbouring offsets are one packed FADD2 (sub.f32x2, round-to-nearest per element,
        // no FTZ: bit-identical to two FADDs); the sum itself stays the reference's sequential fp32 chain (:38-44).
#pragma unroll
        for (int q = 0; q < (P4 + 7) / 8; ++q) {
            if (kSkipGroups && ((um[(8 * q) >> 5] >> ((8 * q) & 31)) & 0xffu) == 0u) continue;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int g = 2 * q + h;
                if (g >= 1 && g < P4 / 4) {
                    float a4[NK][4];
#pragma unroll
                    for (int k = 0; k < NK; ++k) {
                        const float4 av = lds_v4(a_addr[k] + 16 * g);
                        a4[k][0] = av.x; a4[k][1] = av.y; a4[k][2] = av.z; a4[k][3] = av.w;
                    }
#pragma unroll
                    for (int u = 0; u < 4; u += 2) {
                        const int i = 4 * g + u;
                        const bool v0 = i >= 5 && i < P, v1 = i + 1 >= 5 && i + 1 < P;
#pragma unroll
                        for (int k = 0; k < NK; ++k)
#pragma unroll
                            for (int c = 0; c < CPT; ++c) {
                                float t0 = 0.0f, t1 = 0.0f;
                                if (kPackedSub && v0 && v1) {
                                    fsub2(a4[k][u], a4[k][u + 1], x[c][v0 ? i - 5 : 0], x[c][v1 ? i - 4 : 0], t0, t1);
                                } else {
                                    if (v0) t0 = __fsub_rn(a4[k][u], x[c][v0 ? i - 5 : 0]);
                                    if (v1) t1 = __fsub_rn(a4[k][u + 1], x[c][v1 ? i - 4 : 0]);
                                }
                                if (v0 && (m[k][c][i >> 5] & (1u << (i & 31)))) dist[k][c] = __fadd_rn(dist[k][c], fabsf(t0));
                                if (v1 && (m[k][c][(i + 1) >> 5] & (1u << ((i + 1) & 31))))
                                    dist[k][c] = __fadd_rn(dist[k][c], fabsf(t1));
                            }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < NK; ++k)
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const bool hit = act[k][c] && (dist[k][c] < __fmul_rn(p.thr, (float)len[k][c]));  // :46
            hit_out[k][c] = hit;
            if (hit || myK[c] == wk[k]) par[c] = (uint32_t)(n_base + k + 1);                  // :127,:129
        }
}

// The planned first batch knows ALL of its kept lanes before a column is touched, so they are evaluated in ONE pass over
// the registers: NK independent fp32 chains per proposal (each still the reference's ascending sequential sum, :38-44)
// instead of NK passes with one chain each.  A pass with a single chain is latency bound -- 72 dependent FADDs, 4 warps
// per scheduler (ncu: issue slots 53 % busy, top stall "wait") -- NK chains hide that latency and the per-pass set-up
// and tail are paid once.  Row word i of kept lane k enters proposal c's sum iff bit i of (ma_k & mb_c) is set: the
// pair's range [max(sa,sb), min(ea,eb)] is the intersection of the two lanes' own ranges.  That identity needs
// max(sa,sb) >= 0 (below, header words or the wrapped unsigned-char counter come into play, :38): if any active pair
// of the warp has a negative start, the warp takes the exact one-lane-per-pass evaluator instead (rare).
// Lanes that are not "active" for a kept lane (ranked before it, empty range) accumulate a value nobody reads.
template <int NOFF, int CPT, int NK, typename HdrFn>
__device__ __forceinline__ void freg_eval_multi(const FusedParams &p, long long f, const unsigned char *const (&hdr)[NK],
                                                int nk, const bool (&live)[CPT], const u64 (&myK)[CPT],
                                                const int (&st)[CPT], const int (&en)[CPT],
                                                const uint32_t (&mb)[CPT][(5 + NOFF + 31) / 32], const float (&x)[CPT][NOFF],
                                                HdrFn my_hdr, uint32_t (&par)[CPT], long long n_base) {
    constexpr int P = 5 + NOFF, MW = (P + 31) / 32, P4 = (P + 3) & ~3;
    u64 wk[NK];
    int sa[NK], ea[NK];
    uint32_t a_addr[NK];
    bool rare = false;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
        const uint4 wh = *reinterpret_cast<const uint4 *>(hdr[k]);
        wk[k] = k < nk ? (((u64)wh.x << 32) | wh.y) : kNone64;   // a padding lane activates nobody
        sa[k] = (int)wh.z;
        ea[k] = (int)wh.w;
        a_addr[k] = smem_u32(hdr[k] + kHdr);
#pragma unroll
        for (int c = 0; c < CPT; ++c)
            rare |= live[c] && (myK[c] > wk[k]) && (min(ea[k], en[c]) >= max(sa[k], st[c])) && (max(sa[k], st[c]) < 0);
    }
    if (__any_sync(0xffffffffu, rare)) {
        for (int k = 0; k < nk; ++k) {
            const unsigned char *const hh[1] = {hdr[k]};
            bool hit[1][CPT];
            freg_eval<NOFF, CPT, 1>(p, f, hh, live, myK, st, en, mb, x, my_hdr, par, hit, n_base + k);
        }
        return;
    }
    float dist[NK][CPT];
    constexpr bool kSkip = kSkipGroups && CPT >= 2;
#pragma unroll
    for (int k = 0; k < NK; ++k)
#pragma unroll
        for (int c = 0; c < CPT; ++c) dist[k][c] = 0.0f;
    if constexpr (kSkip) {
#pragma unroll
        for (int w = 0; w < MW; ++w) {
            uint32_t m[NK][CPT];   // pair masks of this 32-word span
            uint32_t um[NK];       // the kept lanes' own masks, through REDUX: a uniform register, so the skips below are
                                   // warp-uniform branches (no reconvergence code)
#pragma unroll
            for (int k = 0; k < NK; ++k) {
                const uint32_t ma = *reinterpret_cast<const uint32_t *>(hdr[k] + 16 + 4 * w);
                um[k] = kSkip ? __reduce_or_sync(0xffffffffu, ma) : 0xffffffffu;
#pragma unroll
                for (int c = 0; c < CPT; ++c) m[k][c] = ma & mb[c][w];
            }
            // groups of 8 row words (two LDS.128 of a kept row); a group outside kept lane k's own range adds nothing to any
            // of its pairs (the pair mask is a subset of the lane's mask) and is skipped by the whole warp.  Only with two
            // proposals per thread (four chains: measured +9 % at n_off 36); with one proposal per thread the branches cost
            // more in exposed LDS latency than the skipped work saves (measured -13 % at n_off 72)
#pragma unroll
            for (int q = 4 * w; q < 4 * w + 4; ++q) {
#pragma unroll
                for (int k = 0; k < NK; ++k) {
                    if (kSkip && ((um[k] >> ((8 * q) & 31)) & 0xffu) == 0u) continue;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int g = 2 * q + h;
                        if (g >= 1 && g < P4 / 4) {
                            const float4 av = lds_v4(a_addr[k] + 16 * g);
                            const float a4[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
                            for (int u = 0; u < 4; u += 2) {
                                const int i = 4 * g + u;
                                const bool v0 = i >= 5 && i < P, v1 = i + 1 >= 5 && i + 1 < P;
#pragma unroll
                                for (int c = 0; c < CPT; ++c) {
                                    float t0 = 0.0f, t1 = 0.0f;
                                    if (v0 && v1) {
                                        fsub2(a4[u], a4[u + 1], x[c][v0 ? i - 5 : 0], x[c][v1 ? i - 4 : 0], t0, t1);
                                    } else {
                                        if (v0) t0 = __fsub_rn(a4[u], x[c][v0 ? i - 5 : 0]);
                                        if (v1) t1 = __fsub_rn(a4[u + 1], x[c][v1 ? i - 4 : 0]);
                                    }
                                    if (v0 && (m[k][c] & (1u << (i & 31)))) dist[k][c] = __fadd_rn(dist[k][c], fabsf(t0));
                                    if (v1 && (m[k][c] & (1u << ((i + 1) & 31)))) dist[k][c] = __fadd_rn(dist[k][c], fabsf(t1));
                                }
                            }
                        }
                    }
                }
            }
        }
    } else {
#pragma unroll
        for (int w = 0; w < MW; ++w) {
            uint32_t m[NK][CPT];   // pair masks of this 32-word span
#pragma unroll
            for (int k = 0; k < NK; ++k) {
                const uint32_t ma = *reinterpret_cast<const uint32_t *>(hdr[k] + 16 + 4 * w);
#pragma unroll
                for (int c = 0; c < CPT; ++c) m[k][c] = ma & mb[c][w];
            }
#pragma unroll
            for (int g = 8 * w; g < 8 * w + 8; ++g) {
                if (g >= 1 && g < P4 / 4) {
#pragma unroll
                    for (int k = 0; k < NK; ++k) {
                        const float4 av = lds_v4(a_addr[k] + 16 * g);
                        const float a4[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
                        for (int u = 0; u < 4; u += 2) {
                            const int i = 4 * g + u;
                            const bool v0 = i >= 5 && i < P, v1 = i + 1 >= 5 && i + 1 < P;
#pragma unroll
                            for (int c = 0; c < CPT; ++c) {
                                float t0 = 0.0f, t1 = 0.0f;
                                if (v0 && v1) {
                                    fsub2(a4[u], a4[u + 1], x[c][v0 ? i - 5 : 0], x[c][v1 ? i - 4 : 0], t0, t1);
                                } else {
                                    if (v0) t0 = __fsub_rn(a4[u], x[c][v0 ? i - 5 : 0]);
                                    if (v1) t1 = __fsub_rn(a4[u + 1], x[c][v1 ? i - 4 : 0]);
                                }
                                if (v0 && (m[k][c] & (1u << (i & 31)))) dist[k][c] = __fadd_rn(dist[k][c], fabsf(t0));
                                if (v1 && (m[k][c] & (1u << ((i + 1) & 31)))) dist[k][c] = __fadd_rn(dist[k][c], fabsf(t1));
                            }
                        }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < NK; ++k)      // in keep order: the last kept lane that covers a proposal wins (:127)
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const int start = max(sa[k], st[c]), end = min(ea[k], en[c]);                  // :31,:34
            const bool act = live[c] && (myK[c] > wk[k]) && (end >= start);                 // :36
            const int len = (int)((uint32_t)end - (uint32_t)start + 1u);
            const bool hit = act && (dist[k][c] < __fmul_rn(p.thr, (float)len));            // :46
            if (k < nk && (hit || myK[c] == wk[k])) par[c] = (uint32_t)(n_base + k + 1);    // :127,:129
        }
}

#define PHNMS_TRACE(tag)                                                                      \
    do {                                                                                      \
        if (kTrace && p.trace && p.trace_len > 0 && blockIdx.x == 0 && tid == 0 && tcount + 1 < p.trace_len) { \
            p.trace[tcount++] = (long long)(tag);                                             \
            p.trace[tcount++] = clock64();                                                    \
        }                                                                                     \
    } while (0)

template <int NOFF, int CPT, bool kTrace, bool kDyn>
__global__ void __launch_bounds__(512, 1) phnms_freg_kernel(const FusedParams p, const FregLayout L) {
    constexpr int P = 5 + NOFF;
    constexpr int MW = (P + 31) / 32;   // in-range bitmask words (<= 3)
    constexpr int P4 = (P + 3) & ~3;
    constexpr int SLOT = kHdr + 4 * P4;
    constexpr int M = kCand;

    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    const int csize = p.csize;
    const uint32_t rank = csize > 1 ? cluster_ctarank() : 0u;

    u64 *wtop = reinterpret_cast<u64 *>(smem + L.off_wtop);
    unsigned char *bh = smem + L.off_bh;                                               // [32][kHdr]
    uint32_t *lcount_p = reinterpret_cast<uint32_t *>(smem + L.off_bh + 32 * kHdr);    // [1]
    uint32_t *cdead = lcount_p + 4;                                                    // [3][32]
    float *bit_key = reinterpret_cast<float *>(smem + L.off_bit);
    int *bit_val = reinterpret_cast<int *>(smem + L.off_bit + 128);
    int *bit_ok = reinterpret_cast<int *>(smem + L.off_bit + 256);
    unsigned char *slots = smem + L.off_slots;
    unsigned char *pslots = smem + L.off_pslots;
    unsigned char *rows_buf = smem + L.off_rows;
    float *sbuf = reinterpret_cast<float *>(smem + L.off_sbuf);
    const int sbuf_stride = round_up(p.rpc * 4, 16) / 4;
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

    // ---- frames are CLAIMED, not statically assigned: a global counter hands out frame indices, so a cluster that starts
    // late (its SMs were busy with another kernel, e.g. a concurrent NCCL collective) simply takes fewer frames instead
    // of leaving a second wave.  Rank 0 of the cluster claims two frames ahead and broadcasts the index to its peers with
    // st.async (completing on their claim mbarrier); a relaxed split cluster barrier (arrive at the top of a frame, wait
    // at its end) keeps the CTAs of a cluster within one frame of each other so the two claim slots can be reused.
    // The split barrier costs ~3 % when nothing disturbs the kernel, so clusters of more than one CTA default to the static
    // interleaved assignment (claim_ctr == nullptr); single-CTA frames claim by default (no cluster barrier needed).
    constexpr bool dyn = kDyn;   // compile-time: the static kernel carries none of the claim machinery
    const uint32_t bar_claim0 = bar_load + 24;
    volatile long long *claim_slot = reinterpret_cast<volatile long long *>(smem + 40);
    const long long cl = blockIdx.x / csize, ncl = gridDim.x / csize;
    if (dyn && rank == 0 && tid == 0) {
        const long long c0 = (long long)atomicAdd(p.claim_ctr, 2ull);   // this cluster's first two frames
        if (csize == 1) {
            claim_slot[0] = c0;
        } else {
            for (int d = 0; d < csize; ++d)
                asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(mapa_u32(smem_u32(smem + 40), (uint32_t)d)), "l"(c0) : "memory");
        }
    }
    if (dyn) {
        if (csize > 1) {
            cluster_arrive_release();
            cluster_wait_acquire();
        } else {
            __syncthreads();
        }
    }
    // Resume mode (static schedule only): the frames come from a device-side list -- `fi` walks the list, `f` is the frame.
    const long long Fn = p.frame_list ? (long long)*p.frame_count : p.F;
    auto frame_at = [&](long long i) -> long long { return (p.frame_list && i < Fn) ? (long long)p.frame_list[i] : i; };
    long long fi = dyn ? claim_slot[0] : cl, fi_next = dyn ? fi + 1 : cl + ncl;
    __syncthreads();   // everyone has read slot 0 before a claim may overwrite it
    if (tid == 0) {
        mbar_init(bar_claim0, 1);
        mbar_init(bar_claim0 + 8, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (csize > 1) {   // claim barriers are initialised cluster-wide before the first claim is sent
        cluster_arrive_release();
        cluster_wait_acquire();
    }

    uint32_t load_phase = 0, round_ctr = 0, iter = 0;
    int tcount = 0;
    (void)tcount;
    // slab geometry: identical for every frame in the common case (see Slab), then kept in shared memory
    const bool geo_uniform = slab_is_uniform(p, P);
    // scores ride on the TMA as well when every CTA's slice of every frame is 16-byte aligned and a multiple of 16 bytes
    // up to the end of the frame (T >= 96 so that thread 64 exists)
    const bool sbulk = geo_uniform && ((((uintptr_t)p.scores) | (uintptr_t)(p.N * 4) | (uintptr_t)(p.rpc * 4)) & 15u) == 0u && T >= 96;
    const unsigned char *geo = smem + 64;
    if (geo_uniform && tid == 0) {
        const Slab s0 = slab_compute(p, 0, rank, P, T);
        *reinterpret_cast<int4 *>(smem + 64) = make_int4(s0.nv, s0.r0, s0.nloc, s0.head);
        *reinterpret_cast<int4 *>(smem + 80) = make_int4((int)s0.total, (int)s0.chunk, s0.tail0, s0.ntail);
    }
    __syncthreads();
    if (fi < Fn) {
        const long long f0 = frame_at(fi);
        request_slab(p, f0, slab_geometry(p, f0, rank, P, T, geo_uniform, geo), rows_buf, bar_load, tid, T, P, pslots,
                     (uint32_t)(p.topm_count * SLOT), sbuf, sbulk);
    }
    // Threads beyond this CTA's rows ("spare lanes") hold register copies of the batch's candidates, so that every CTA
    // can tell -- without talking to its peers -- which candidates an earlier winner of the same batch suppressed.
    const int lcap = min(T * CPT - p.rpc, 31);

    for (; fi < Fn; fi = fi_next, ++iter) {
        const long long f = frame_at(fi);
        PHNMS_TRACE(1);  // frame start
        const uint32_t fpar = iter & 1u;   // parity of the double-buffered candidate block / scores
        if (dyn && csize > 1) cluster_arrive_relaxed();
        long long my_claim = 0;
        if (dyn && rank == 0 && tid == 0) my_claim = (long long)atomicAdd(p.claim_ctr, 1ull);   // frame of iteration iter + 2
        // ---- staging -> registers -------------------------------------------------------------------------------
        const Slab cur = slab_geometry(p, f, rank, P, T, geo_uniform, geo);
        if (cur.total != 0u || p.topm != nullptr || (sbulk && cur.nloc > 0)) {
            if (kTrace && p.trace_len < 0) mbar_wait_watch(bar_load, load_phase, p.trace, 1, f, round_ctr);
            else mbar_wait(bar_load, load_phase);
            load_phase ^= 1u;
        }
        cp_async_wait_all();   // my own score copies (each thread reads back only what it copied itself)
        __syncthreads();
        // Dead flags of the first batch.  Reset only here, between the frame's two barriers: before the first one a slower
        // warp can still be inside the previous frame's last round (reading or setting these flags).
        if (warp == 0) cdead[2 * 32 + lane] = 0u;
        PHNMS_TRACE(2);  // slab landed
        const float *rows = reinterpret_cast<const float *>(rows_buf + 16 - cur.head);
        const int nv = cur.nv, r0 = cur.r0, nloc = cur.nloc;
        float x[CPT][NOFF];   // the offsets only; the 5 header words are re-read on the rare paths that need them
        bool real[CPT], virt[CPT];
        uint32_t key[CPT], par[CPT], mb[CPT][MW];
        int st[CPT], en[CPT];
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const int col = c * T + tid;
            real[c] = col < nloc;
            virt[c] = false;
            const float *row = rows + (size_t)(real[c] ? col : 0) * P;
#pragma unroll
            for (int i = 0; i < NOFF; ++i) x[c][i] = row[5 + i];
            st[c] = lane_start(row[2], NOFF);            // nms_kernel.cu:29-30
            en[c] = lane_end(row[4], st[c], NOFF);       // :32-34
            key[c] = key_desc(real[c] ? sbuf[fpar * sbuf_stride + col] : 0.0f, p.sort_model == 1);
            par[c] = 0u;
            range_mask<MW>(st[c], en[c], mb[c]);
        }
        const bool bitonic = (p.sort_model == 0) && nv <= 32 && nv >= 2;  // torch: unstable bitonic network (n <= 32)
        if (bitonic && warp == 0 && nloc > 0) {  // rank 0 holds the whole frame (rows_per_cta >= 32), column == lane
            bit_ok[lane] = lane < nv;
            bit_key[lane] = lane < nv ? sbuf[fpar * sbuf_stride + lane] : 0.0f;
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
            int mypos = 0;  // sorted position -> rank key of the proposal that landed there
            for (int q = 0; q < 32; ++q)
                if (bit_val[q] == lane && q < nv) mypos = q;
            key[0] = (uint32_t)mypos;
        }
        __syncthreads();  // every row is in registers: the staging buffer is free
        PHNMS_TRACE(3);  // rows in registers

        // ---- request the next frame now; it lands while this frame's rounds run ----------------------------------
        if (!dyn) {
            fi_next = fi + ncl;
        } else if (iter > 0) {   // the frame after this one: claimed during the previous iteration, sent in its middle
            const uint32_t cp_ = (iter - 1u) & 1u;
            if (csize > 1) mbar_wait(bar_claim0 + 8u * cp_, ((iter - 1u) >> 1) & 1u);
            fi_next = claim_slot[cp_];
        }
        const bool more = fi_next < Fn;   // (uniform over the cluster) another iteration follows
        if (dyn && more && csize > 1 && tid == 0) mbar_arrive_expect_tx(bar_claim0 + 8u * (iter & 1u), 8u);
        if (more) {
            const long long f_next = frame_at(fi_next);
            request_slab(p, f_next, slab_geometry(p, f_next, rank, P, T, geo_uniform, geo), rows_buf, bar_load, tid, T, P,
                         pslots + (size_t)(fpar ^ 1u) * kTopM * L.slot_stride, (uint32_t)(p.topm_count * SLOT),
                         sbuf + (fpar ^ 1u) * sbuf_stride, sbulk);
        }
        PHNMS_TRACE(4);  // next slab requested
        // hand the claimed frame index to the cluster; peers read it after the second barrier of their next frame
        // (sent here, mid-frame, so that it is long there when they need it)
        if (dyn && csize > 1) cluster_wait_plain();   // every peer has started this frame: slot (iter & 1) was consumed
        if (dyn && more && rank == 0 && tid == 0) {
            if (csize == 1) {
                claim_slot[iter & 1u] = my_claim;
            } else {
                const uint32_t slot_a = smem_u32(smem + 40) + 8u * (iter & 1u), bar_a = bar_claim0 + 8u * (iter & 1u);
                for (int d = 0; d < csize; ++d)
                    st_async_b64(mapa_u32(slot_a, (uint32_t)d), (unsigned long long)my_claim, mapa_u32(bar_a, (uint32_t)d));
            }
        }

        u64 myK[CPT];
#pragma unroll
        for (int c = 0; c < CPT; ++c)
            myK[c] = real[c] ? (((u64)key[c] << 32) | (uint32_t)(r0 + c * T + tid)) : kNone64;

        // ---- greedy rounds (nms_collect, :111-136), in batches.  Batch 0 uses the frame's kTopM best-ranked proposals
        // found by phnms_topm_kernel (every CTA fetched the records and rows itself: no selection, no exchange).  If
        // more lanes are needed, fallback batches follow: every CTA publishes its best kCand alive lanes; the merged
        // list is exact up to the smallest "last published key" (an unpublished alive lane of CTA d ranks after d's
        // last published one) and that guaranteed prefix is consumed with no further communication.
        long long n = 0;
        bool frame_done = false;

        // ---- batch 0, planned: the candidate block carries, for every candidate, which later candidates it suppresses
        // (computed by phnms_topm_kernel).  The greedy scan over the candidates is therefore known before a single
        // column is touched: no selection, no exchange, no barrier -- the kept candidates are evaluated two per pass.
        if (p.topm != nullptr) {
            const unsigned char *csl = pslots + (size_t)fpar * kTopM * L.slot_stride;
            // aux of slot 0: number of valid candidates; aux of slot 1: the kept set of the greedy scan over the candidates,
            // done once per frame by phnms_topm_kernel (which also stored their indices to keep[f, 0 ..], :118)
            const int nc = min((int)(reinterpret_cast<const uint32_t *>(csl)[7] & 0xffffu), p.topm_count);
            uint32_t kept = reinterpret_cast<const uint32_t *>(csl + L.slot_stride)[7] & 0xffffu;
            if (nc < 2) kept = (uint32_t)nc;   // (a frame of one proposal has no valid slot 1)
            auto real_hdr = [&](int c) { return p.props + ((size_t)f * p.N + (uint32_t)myK[c]) * P; };
            while (kept) {
                constexpr int NKP = CPT == 1 ? kPlanLanes : kPlanLanes2;
                const unsigned char *hh[NKP];
                int cnt = 0;
#pragma unroll
                for (int k = 0; k < NKP; ++k) {
                    const int i = kept ? __ffs(kept) - 1 : -1;
                    if (i >= 0) {
                        kept &= kept - 1u;
                        ++cnt;
                    }
                    hh[k] = csl + (size_t)max(i, 0) * L.slot_stride;   // padding lanes point at a valid slot and are ignored
                }
                if (NKP == 1) {
                    const unsigned char *const h1[1] = {hh[0]};
                    bool hit[1][CPT];
                    freg_eval<NOFF, CPT, 1>(p, f, h1, real, myK, st, en, mb, x, real_hdr, par, hit, n);
                } else {
                    const unsigned char *const (&hc)[NKP] = hh;
                    freg_eval_multi<NOFF, CPT, NKP>(p, f, hc, cnt, real, myK, st, en, mb, x, real_hdr, par, n);
                }
                n += cnt;
            }
            PHNMS_TRACE(11);  // planned batch done
            // done when top_k lanes are kept, or when every lane of the frame was a candidate
            if (n == p.top_k || nc >= nv) frame_done = true;
        }

        // ---- fallback batches (needed when the candidates run out before top_k lanes are kept): every CTA publishes its
        // best kCand alive lanes; the merged list is exact up to the smallest "last published key" (an unpublished alive
        // lane of CTA d ranks after d's last published one) and that guaranteed prefix is consumed with no further
        // communication -- spare lanes hold register copies of the candidates and flag the ones a winner suppresses.
        while (!frame_done) {
            const unsigned char *hb;   // headers of this batch in rank order
            int hs;                    // header stride
            const unsigned char *sl;   // slots that hold the rows
            uint32_t dead_row;
            int lcount;
            {
                const uint32_t par_bit = round_ctr & 1u, xphase = (round_ctr >> 1) & 1u;
                ++round_ctr;
                dead_row = par_bit;
                // 1. warp-local best M alive lanes
                u64 wbest[M];
                bool taken[CPT];
#pragma unroll
                for (int c = 0; c < CPT; ++c) taken[c] = false;
#pragma unroll
                for (int i = 0; i < M; ++i) {
                    u64 cand = kNone64;
#pragma unroll
                    for (int c = 0; c < CPT; ++c)
                        if (real[c] && par[c] == 0u && !taken[c]) cand = min(cand, myK[c]);
                    const u64 b = warp_min_u64(cand);
#pragma unroll
                    for (int c = 0; c < CPT; ++c)
                        if (b != kNone64 && myK[c] == b) taken[c] = true;
                    wbest[i] = b;
                }
                {
                    u64 v = wbest[0];
#pragma unroll
                    for (int i = 1; i < M; ++i)
                        if (lane == i) v = wbest[i];
                    if (lane < M) wtop[(par_bit * 32 + warp) * M + lane] = v;
                }
                __syncthreads();
                PHNMS_TRACE(5);  // warp-level candidates
                // 2. CTA-level best M (every warp redundantly; nwarps * M <= 64 entries)
                u64 ctop[M];
                {
                    const int nE = nwarps * M;
                    const u64 *wt = wtop + (size_t)par_bit * 32 * M;
                    u64 e0 = lane < nE ? wt[lane] : kNone64, e1 = lane + 32 < nE ? wt[lane + 32] : kNone64;
#pragma unroll
                    for (int i = 0; i < M; ++i) {
                        const u64 b = warp_min_u64(min(e0, e1));
                        ctop[i] = b;
                        if (b != kNone64) {
                            if (e0 == b) e0 = kNone64;
                            if (e1 == b) e1 = kNone64;
                        }
                    }
                }
                PHNMS_TRACE(6);  // CTA-level candidates
                // 3. publish candidate i {header, row} into slot[par][rank][i] of every CTA
                const uint32_t slot_base = smem_u32(slots + (size_t)par_bit * csize * M * L.slot_stride);
                const uint32_t bar_x = bar_x0 + 8u * par_bit;
                if (csize > 1 && tid == 0) mbar_arrive_expect_tx(bar_x, (uint32_t)(csize * M * SLOT));
#pragma unroll
                for (int c = 0; c < CPT; ++c) {
                    int mine = -1;
#pragma unroll
                    for (int i = 0; i < M; ++i)
                        if (real[c] && ctop[i] != kNone64 && myK[c] == ctop[i]) mine = i;
                    if (mine >= 0) {
                        // the published row carries its 5 header words too (re-read from global memory: this is the
                        // rare second batch of a frame)
                        float hdr[5];
                        const float *grow = p.props + ((size_t)f * p.N + (uint32_t)myK[c]) * P;
#pragma unroll
                        for (int i = 0; i < 5; ++i) hdr[i] = grow[i];
#define ROWW(i) ((i) < 5 ? __float_as_uint(hdr[(i) < 5 ? (i) : 0]) : ((i) < P ? __float_as_uint(x[c][(i) >= 5 && (i) < P ? (i) - 5 : 0]) : 0u))
                        const uint32_t myslot = slot_base + (uint32_t)(((int)rank * M + mine) * L.slot_stride);
                        const uint32_t khi = (uint32_t)(myK[c] >> 32), klo = (uint32_t)myK[c];
                        const uint32_t m0 = mb[c][0], m1 = MW > 1 ? mb[c][MW > 1 ? 1 : 0] : 0u, m2 = MW > 2 ? mb[c][MW > 2 ? 2 : 0] : 0u;
                        if (csize == 1) {
                            sts_v4(myslot, khi, klo, (uint32_t)st[c], (uint32_t)en[c]);
                            sts_v4(myslot + 16, m0, m1, m2, 0u);
#pragma unroll
                            for (int g = 0; g < P4 / 4; ++g)
                                sts_v4(myslot + kHdr + 16 * g, ROWW(4 * g), ROWW(4 * g + 1), ROWW(4 * g + 2), ROWW(4 * g + 3));
                        } else {
                            for (int d = 0; d < csize; ++d) {
                                const uint32_t dst = mapa_u32(myslot, (uint32_t)d), dbar = mapa_u32(bar_x, (uint32_t)d);
                                st_async_v4(dst, khi, klo, (uint32_t)st[c], (uint32_t)en[c], dbar);
                                st_async_v4(dst + 16, m0, m1, m2, 0u, dbar);
#pragma unroll
                                for (int g = 0; g < P4 / 4; ++g)
                                    st_async_v4(dst + kHdr + 16 * g, ROWW(4 * g), ROWW(4 * g + 1), ROWW(4 * g + 2), ROWW(4 * g + 3), dbar);
                            }
                        }
#undef ROWW
                    }
                }
                if (tid < M) {  // fewer than M alive lanes here: empty slots of the same size
                    bool empty = false;
#pragma unroll
                    for (int i = 0; i < M; ++i)
                        if (tid == i && ctop[i] == kNone64) empty = true;
                    if (empty) {
                        const uint32_t myslot = slot_base + (uint32_t)(((int)rank * M + tid) * L.slot_stride);
                        if (csize == 1) {
                            sts_v4(myslot, 0xffffffffu, 0xffffffffu, 0u, 0u);
                        } else {
                            for (int d = 0; d < csize; ++d) {
                                const uint32_t dst = mapa_u32(myslot, (uint32_t)d), dbar = mapa_u32(bar_x, (uint32_t)d);
                                st_async_v4(dst, 0xffffffffu, 0xffffffffu, 0u, 0u, dbar);
                                for (int g = 1; g < SLOT / 16; ++g) st_async_v4(dst + 16 * g, 0u, 0u, 0u, 0u, dbar);
                            }
                        }
                    }
                }
                PHNMS_TRACE(7);  // published
                if (csize > 1) {
                    if (kTrace && p.trace_len < 0) mbar_wait_watch(bar_x, xphase, p.trace, 2, f, round_ctr);
                    else mbar_wait(bar_x, xphase);
                } else {
                    __syncthreads();
                }
                PHNMS_TRACE(8);  // exchange complete

                // 4. merged guaranteed prefix (warp 0): rank every published candidate, keep those not above the bound,
                //    and copy their headers, in rank order, into the compact array the rounds read
                sl = slots + (size_t)par_bit * csize * M * L.slot_stride;
                hb = bh;
                hs = kHdr;
                if (warp == 0) {
                    const int nS = csize * M;
                    u64 bound = kNone64;
                    for (int d = 0; d < csize; ++d) {
                        const uint2 h = *reinterpret_cast<const uint2 *>(sl + (size_t)(d * M + M - 1) * L.slot_stride);
                        bound = min(bound, ((u64)h.x << 32) | h.y);   // an incomplete list (empty last slot) bounds nothing
                    }
                    int nvalid = 0;
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const int e = lane + 32 * half;
                        u64 k = kNone64;
                        uint4 h0 = make_uint4(0xffffffffu, 0xffffffffu, 0u, 0u), h1 = make_uint4(0u, 0u, 0u, 0u);
                        if (e < nS) {
                            h0 = *reinterpret_cast<const uint4 *>(sl + (size_t)e * L.slot_stride);
                            h1 = *reinterpret_cast<const uint4 *>(sl + (size_t)e * L.slot_stride + 16);
                            k = ((u64)h0.x << 32) | h0.y;
                        }
                        int rk = 0;
                        for (int q = 0; q < nS; ++q) {
                            const uint2 hq = *reinterpret_cast<const uint2 *>(sl + (size_t)q * L.slot_stride);
                            rk += (((u64)hq.x << 32) | hq.y) < k;
                        }
                        const bool ok = k != kNone64 && k <= bound;
                        if (ok && rk <= lcap) {
                            h1.w = (uint32_t)e;   // aux: the slot that holds the row
                            *reinterpret_cast<uint4 *>(bh + (size_t)rk * kHdr) = h0;
                            *reinterpret_cast<uint4 *>(bh + (size_t)rk * kHdr + 16) = h1;
                        }
                        nvalid += __popc(__ballot_sync(0xffffffffu, ok));
                    }
                    if (lane == 0) *lcount_p = (uint32_t)min(nvalid, 1 + lcap);
                    cdead[dead_row * 32 + lane] = 0u;
                }
                __syncthreads();
                PHNMS_TRACE(9);  // merged list ready
                lcount = (int)*lcount_p;
            }
            if (lcount == 0) break;  // every lane is kept or removed

            // 5. spare lanes take register copies of candidates 1 .. lcount-1
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const int v = c * T + tid - p.rpc;
                if (v >= 0) {
                    virt[c] = 1 + v < lcount;
                    myK[c] = kNone64;
                    if (virt[c]) {
                        const unsigned char *h = hb + (size_t)(1 + v) * hs;
                        const uint4 h0 = *reinterpret_cast<const uint4 *>(h);
                        const uint4 h1 = *reinterpret_cast<const uint4 *>(h + 16);
                        myK[c] = ((u64)h0.x << 32) | h0.y;
                        st[c] = (int)h0.z;
                        en[c] = (int)h0.w;
                        mb[c][0] = h1.x;
                        if (MW > 1) mb[c][MW > 1 ? 1 : 0] = h1.y;
                        if (MW > 2) mb[c][MW > 2 ? 2 : 0] = h1.z;
                        const float *row = reinterpret_cast<const float *>(sl + (size_t)h1.w * L.slot_stride + kHdr);
#pragma unroll
                        for (int i = 0; i < NOFF; ++i) x[c][i] = row[5 + i];
                        par[c] = 0u;
                    }
                }
            }
            PHNMS_TRACE(10);  // spare lanes loaded

            // 6. one round per alive candidate of the prefix, in rank order
            auto any_hdr = [&](int c) -> const float * {
                if (real[c]) return p.props + ((size_t)f * p.N + (uint32_t)myK[c]) * P;
                const int v = c * T + tid - p.rpc;
                return reinterpret_cast<const float *>(
                    sl + (size_t)reinterpret_cast<const uint32_t *>(hb + (size_t)(1 + v) * hs)[7] * L.slot_stride + kHdr);
            };
            bool live[CPT];
#pragma unroll
            for (int c = 0; c < CPT; ++c) live[c] = real[c] || virt[c];
            for (int j = 0; j < lcount; ++j) {
                if (j > 0 && cdead[dead_row * 32 + j] != 0u) continue;  // suppressed by an earlier winner of this batch
                const unsigned char *h = hb + (size_t)j * hs;
                // the compact header points at the slot that holds the row: build {header, row} view of that slot
                const unsigned char *slot = sl + (size_t)reinterpret_cast<const uint32_t *>(h)[7] * L.slot_stride;
                if (rank == 0 && tid == 0) p.keep[(size_t)f * p.N + n] = (long long)reinterpret_cast<const uint32_t *>(h)[1];
                const unsigned char *const hh[1] = {slot};
                bool hit[1][CPT];
                freg_eval<NOFF, CPT, 1>(p, f, hh, live, myK, st, en, mb, x, any_hdr, par, hit, n);
                // A suppressed candidate is flagged for the LATER round that would have picked it (read after at least
                // one barrier).  The winner's own copy must not touch its flag: slower warps may not have read it yet at
                // the top of THIS round (that race skipped the round in some warps -> hang).
#pragma unroll
                for (int c = 0; c < CPT; ++c)
                    if (virt[c] && hit[0][c]) cdead[dead_row * 32 + 1 + (c * T + tid - p.rpc)] = 1u;
                ++n;
                if (n == p.top_k) {  // :133 (top_k == 0 never stops early)
                    frame_done = true;
                    break;
                }
                __syncthreads();  // candidate-dead flags of this round are visible to the next
            }
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
            if (rank == 0 && tid == 0) {
                const long long num = p.top_k < n ? p.top_k : n;
                p.num_keep[f] = num;  // :142
                if (p.rec.n > 0) {    // the frame's compact record (keep[0 .. n) was written by this thread, or by the kernel before)
                    for (int c = 0; c < p.rec.width - 1; ++c) record_store(p.rec, f, c, c < num ? keep_f[c] : 0ll);
                    record_store(p.rec, f, p.rec.width - 1, num);
                }
            }
        }
        PHNMS_TRACE(13);  // outputs written

    }

    if (csize > 1) {  // no CTA leaves while a peer may still address its shared memory
        cluster_arrive_release();
        cluster_wait_acquire();
    }
}

}  // namespace phnms

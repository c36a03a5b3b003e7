// stream.cuh -- kernel (2)+(3) of the streaming fast path: every proposal of every frame against the frame's kept lanes.
//
// phnms_select_kernel (select.cuh) has already run the greedy scan (nms_collect, libs/ops/csrc/nms_kernel.cu:99-143) and
// knows the kept lanes of each frame.  What is left of the reference's work is the part that touches all the data: the
// mask rows of the kept lanes (nms_kernel, :50-96, of which nms_collect only ever reads the rows of kept lanes, :116-122)
// and the parent stamps (:123-129).  With the kept lanes known up front that is a pure streaming map over proposals:
//
//     parent[j] = 1 + max{ k : kept lane k is ranked before j and devIoU(kept_k, j) }   (or k + 1 if j IS kept lane k)
//
// No cluster, no per-frame barrier, no exchange: the unit of work is an ITEM of 32 x CPT rows owned by one warp (CPT rows per
// thread: 1 at 72 offsets, 2 at 36 -- 72 offset registers per thread either way).
//   * each warp owns a private staging slot in shared memory and feeds it itself: one TMA 1-D bulk copy (UBLKCP) of the
//     item's 16-byte aligned body + 4-byte cp.async (LDGSTS) for <= 3 unaligned words at either end and for the 32 scores,
//     all completing on the warp's own mbarrier (cp.async.mbarrier.arrive.noinc);
//   * as soon as the rows are in REGISTERS (one proposal per thread, 72 offsets) the slot is free and the warp requests its
//     next item, which lands while the current one is evaluated -- 16 warps per SM = 16 independent load streams, ~160 KB
//     in flight per SM, and the phases of different warps overlap instead of running in lock step;
//   * the kept lanes of a frame (block written by the select kernel) are shared by the CTA through a small ring of
//     shared-memory slots with full / empty mbarriers; whichever warp first needs a block claims the request (atomicCAS);
//   * evaluation: freg_eval_multi (fused_reg.cuh) -- NKP kept lanes per pass over the registers, each chain the reference's
//     ascending sequential fp32 sum (:38-44), packed FADD2 subtractions, R2P predicates; negative common starts (header
//     words / the wrapped unsigned-char counter, :38) take the exact one-lane evaluator.
//   * an OPEN frame (select hit its draw cap) that still has a proposal no kept lane covers is appended to the resume list.
// Roofline: HBM (rows + scores read once, parent / keep padding written once; the kept blocks are L2 hits).
#pragma once
#include "common.cuh"
#include "fused_reg.cuh"
#include "select.cuh"

namespace phnms {

constexpr int kStreamMaxWarps = 16;
constexpr int kStreamMaxRing = 32;

struct StreamParams {
    const float *props;
    const float *scores;
    const int32_t *n_valid;
    long long *keep;
    long long *parent;
    const unsigned char *blocks;
    int block_bytes;
    int *flags;
    unsigned int *ctrs;
    int *list;
    long long F;
    int N, top_k, sort_model;
    float thr;
    int ipf;    // items per frame: ceil(N / rows per item), rows per item = 32 x rows per thread
    int nseg;   // a frame is cut into nseg units of ips item slots (nseg > 1 only when there are fewer frames than CTAs)
    int ips;
    int bundle; // ... or a unit is a bundle of consecutive frames (small frames: so that every warp has an item in every unit)
    int ks;     // kept-block ring slots
    int off_ring, off_slots, slot_bytes, off_bit;
};

struct StreamLayout {
    int off_ring, off_slots, slot_bytes, off_bit, total;
};

// shared memory: [0,128) row mbarriers (one per warp) | [128,384) kfull | [384,640) kempty | [640] next ticket | [644] tickets issued
__host__ __device__ inline StreamLayout stream_layout(int warps, int P, int block_bytes, int ks, int cpt, int bundle) {
    StreamLayout L;
    int o = 768;
    L.off_ring = o;
    o += ks * block_bytes * (bundle > 1 ? bundle : 1);
    o = (o + 127) & ~127;
    L.off_slots = o;
    L.slot_bytes = 128 * cpt + ((16 + 32 * cpt * P * 4 + 16 + 127) & ~127);   // scores | rows (16 B lead for the alignment shift)
    o += warps * L.slot_bytes;
    L.off_bit = o;
    o += warps * 384;
    L.total = o;
    return L;
}

__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// devIoU (nms_kernel.cu:26-48) of NK kept lanes -- consecutive slots starting at shared address s0 -- against this thread's
// CPT proposals, in ONE pass over its registers: NK x CPT independent fp32 chains, each the reference's ascending sequential
// sum (:38-44).  Row word i of kept lane k enters a proposal's sum iff bit i of (ma_k & mb) is set: the pair's range
// [max(sa, sb), min(ea, eb)] is the intersection of the two lanes' own ranges -- which needs max(sa, sb) >= 0 (below, header
// words or the wrapped unsigned-char counter come into play, :38): if any active pair of the warp has a negative start the
// function returns false and the caller takes the exact one-lane evaluator (freg_eval).  All addresses are one register +
// immediates; what stays live across the pass is dist / limit per chain and one word of flags.
template <int NOFF, int NK, int CPT>
__device__ __forceinline__ bool stream_eval(uint32_t s0, int cnt, const bool (&live)[CPT], const u64 (&myK)[CPT],
                                            const int (&st)[CPT], const int (&en)[CPT],
                                            const uint32_t (&mb)[CPT][(5 + NOFF + 31) / 32], const float (&x)[CPT][NOFF],
                                            float thr, uint32_t (&par)[CPT], int k0) {
    constexpr int P = 5 + NOFF, MW = (P + 31) / 32, P4 = (P + 3) & ~3, SLOT = kHdr + 4 * P4;
    float lim[NK][CPT], dist[NK][CPT];
    uint32_t self = 0u;
    bool rare = false;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
        const uint4 wh = lds_u4(s0 + k * SLOT);
        const u64 wk = k < cnt ? (((u64)wh.x << 32) | wh.y) : kNone64;   // a padding lane activates nobody
        const int sa = (int)wh.z, ea = (int)wh.w;
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const int start = max(sa, st[c]), end = min(ea, en[c]);           // :31,:34 (both ends clamped to NOFF-1)
            const bool act = live[c] && (myK[c] > wk) && (end >= start);      // ranked after the kept lane; :36
            rare |= act && (start < 0);
            const int len = (int)((uint32_t)end - (uint32_t)start + 1u);
            lim[k][c] = act ? __fmul_rn(thr, (float)len) : -__int_as_float(0x7f800000);   // :46; -inf: never a hit
            if (myK[c] == wk) self |= 1u << (k * CPT + c);
            dist[k][c] = 0.0f;
        }
    }
    if (__any_sync(0xffffffffu, rare)) return false;
#pragma unroll
    for (int w = 0; w < MW; ++w) {
        uint32_t m[NK][CPT];   // pair masks of this 32-word span
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            const uint32_t ma = lds_u32(s0 + k * SLOT + 16 + 4 * w);
#pragma unroll
            for (int c = 0; c < CPT; ++c) m[k][c] = ma & mb[c][w];
        }
#pragma unroll
        for (int g = 8 * w; g < 8 * w + 8; ++g) {
            if (g >= 1 && g < P4 / 4) {
#pragma unroll
                for (int k = 0; k < NK; ++k) {
                    const float4 av = lds_v4(s0 + k * SLOT + kHdr + 16 * g);
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
#pragma unroll
    for (int k = 0; k < NK; ++k)      // in keep order: the last kept lane that covers a proposal wins (:127)
#pragma unroll
        for (int c = 0; c < CPT; ++c)
            if (k < cnt && ((dist[k][c] < lim[k][c]) || ((self >> (k * CPT + c)) & 1u))) par[c] = (uint32_t)(k0 + k + 1);   // :46,:127,:129
    return true;
}

// GEN = false: the plain big-batch case -- a unit is one whole frame (nseg == 1, bundle == 1) and there is no n_valid; everything
// about an item follows from (us, ci) in a handful of instructions.  GEN = true: segments / bundles / ragged frames.
template <int NOFF, int NKP, bool GEN>
__global__ void __launch_bounds__(kStreamMaxWarps * 32, 1) phnms_stream_kernel(const StreamParams sp) {
    constexpr int P = 5 + NOFF, MW = (P + 31) / 32, P4 = (P + 3) & ~3, SLOT = kHdr + 4 * P4, RPI = 32;   // rows per item
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const uint32_t smem_s = smem_u32(smem);
    const uint32_t bar_rows = smem_s + 8u * warp;
    const uint32_t kfull0 = smem_s + 128u, kempty0 = smem_s + 384u;
    uint32_t *next_req = reinterpret_cast<uint32_t *>(smem + 640);
    const uint32_t ring_s = smem_s + (uint32_t)sp.off_ring;
    const uint32_t slot_s = smem_s + (uint32_t)sp.off_slots + (uint32_t)warp * (uint32_t)sp.slot_bytes;   // 32 scores, then the rows
    const uint32_t ks = (uint32_t)sp.ks, ips = (uint32_t)sp.ips;
    const int nseg = GEN ? sp.nseg : 1, bundle = GEN ? sp.bundle : 1;

    if (threadIdx.x == 0) {
        for (int w = 0; w < nwarps; ++w) mbar_init(smem_s + 8u * w, 33);   // 1 expect_tx arrive + 32 cp.async arrives
        for (uint32_t s = 0; s < ks; ++s) {
            mbar_init(kfull0 + 8u * s, 1);
            mbar_init(kempty0 + 8u * s, ips);   // one arrive per item slot of the unit
        }
        next_req[0] = 0u;
        next_req[1] = 0u;   // `issued`
        fence_mbar_init();
    }
    __syncthreads();

    // Units of this CTA: u = b, b + G, ...  A unit is one frame; or, when there are fewer frames than SMs, one of nseg segments
    // of a frame; or, for small frames, a bundle of consecutive frames (so that every warp has an item in every unit).  The
    // unit's item slots ci = 0 .. ips-1 go round the warps: warp w takes the CTA's item slots w, w + nwarps, ..., tracked
    // incrementally as (us, ci) -- no division on the per-item path unless a unit has fewer item slots than the CTA has warps.
    const long long U = bundle > 1 ? (sp.F + bundle - 1) / bundle : sp.F * nseg;
    const uint32_t b = blockIdx.x, G = gridDim.x;
    const uint32_t nu = b < U ? (uint32_t)((U - b + G - 1) / G) : 0u;
    const int look = (int)ks - 2;
    const uint32_t ring_slot = (uint32_t)sp.block_bytes * (uint32_t)(bundle > 1 ? bundle : 1);
    if (nu == 0u) return;

    struct Item { long long f; int r0, fb; bool valid; };
    auto locate = [&](uint32_t us, uint32_t ci) {
        Item it;
        const long long u = (long long)b + (long long)us * G;
        int c;
        it.fb = 0;
        if (bundle > 1) {          // unit = frames u * bundle ..; item slot ci = (frame within the bundle, item of the frame)
            it.fb = (int)(ci / (uint32_t)sp.ipf);
            c = (int)ci - it.fb * sp.ipf;
            it.f = u * bundle + it.fb;
            it.valid = it.f < sp.F;
            if (!it.valid) it.f = sp.F - 1;
        } else {
            it.f = nseg == 1 ? u : u / nseg;
            const int sg = (int)(u - it.f * nseg);
            c = sg * (int)ips + (int)ci;
            it.valid = c < sp.ipf;
        }
        it.r0 = c * RPI;
        return it;
    };
    auto rows_of = [&](const Item &it) {
        int nv = sp.N;
        if (GEN && sp.n_valid) nv = max(0, min(sp.n_valid[it.f], sp.N));
        return it.valid ? max(0, min(nv - it.r0, RPI)) : 0;
    };
    // the item's rows -> this warp's slot.  Row data keeps its global address modulo 16 (rows are only 4-byte aligned:
    // 308 / 164 bytes), so the aligned body is one bulk copy and at most 3 words at either end are copied singly.
    auto issue = [&](const Item &it, int nrows) {
        if (nrows <= 0) return;
        const float *src = sp.props + ((size_t)it.f * sp.N + it.r0) * P;
        const uintptr_t a0 = (uintptr_t)src;
        const uint32_t bytes = (uint32_t)nrows * (P * 4);
        if (((a0 | bytes) & 15u) == 0u) {   // the usual case: N a multiple of 4 and an aligned tensor
            if (lane == 0) {   // (an L2 evict-first hint on this copy was measured: no gain at N = 1000, -3 % at N = 2048 / 4096)
                mbar_arrive_expect_tx(bar_rows, bytes);
                bulk_g2s(slot_s + 128u, src, bytes, bar_rows);
            }
        } else {
            const uintptr_t b0 = (a0 + 15) & ~(uintptr_t)15, e0 = (a0 + bytes) & ~(uintptr_t)15;
            const uint32_t D = slot_s + 128u + (uint32_t)(a0 & 15);
            if (lane == 0) {
                mbar_arrive_expect_tx(bar_rows, (uint32_t)(e0 - b0));
                bulk_g2s(D + (uint32_t)(b0 - a0), reinterpret_cast<const void *>(b0), (uint32_t)(e0 - b0), bar_rows);
            }
            const int hw = (int)((b0 - a0) >> 2), tw = (int)((a0 + bytes - e0) >> 2), t0 = (int)((e0 - a0) >> 2);
            if (lane >= 1 && lane - 1 < hw) cp_async_4(D + 4u * (uint32_t)(lane - 1), src + (lane - 1));
            if (lane >= 4 && lane - 4 < tw) cp_async_4(D + 4u * (uint32_t)(t0 + lane - 4), src + t0 + (lane - 4));
        }
        if (lane < nrows) cp_async_4(slot_s + 4u * (uint32_t)lane, sp.scores + (size_t)it.f * sp.N + it.r0 + lane);
        cp_async_mbar_arrive_noinc(bar_rows);
    };
    // kept blocks: requested up to `look` units ahead by whichever warp gets there first (lane 0 only).  Requests go out in
    // TICKET ORDER (`issued` counts them): mbarrier waits only tell even from odd phases, so nobody may get two phases ahead
    // of a barrier -- when a unit has fewer item slots than the CTA has warps, different warps work on different units and
    // drift apart; without the ordering a warp far ahead could pass the wait for "the previous user of this ring slot is
    // done" on the strength of the user before that.  A ticket waits for units older than every unit its claimer (or any warp
    // spinning on `issued`) is working on, so the oldest unfinished unit can always finish: no cycle.
    volatile uint32_t *issued = reinterpret_cast<volatile uint32_t *>(smem + 644);
    auto spin_until_issued = [&](uint32_t t) {   // returns once ticket t has been requested
        if (*issued > t) return;
        unsigned long long t0, now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (*issued <= t) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (now - t0 > kMbarTimeoutNs) __trap();
        }
    };
    auto ensure_requested = [&](long long upto) {
        for (;;) {
            const uint32_t t = *reinterpret_cast<volatile uint32_t *>(next_req);
            if ((long long)t > upto || t >= nu) break;
            if (atomicCAS(next_req, t, t + 1u) != t) continue;
            if (t > 0u) spin_until_issued(t - 1u);
            const uint32_t slot = t % ks, use = t / ks;
            if (use > 0u) mbar_wait(kempty0 + 8u * slot, (use - 1u) & 1u);   // every item slot of the previous user is done
            const long long uq = (long long)b + (long long)t * G;
            const long long fq = bundle > 1 ? uq * bundle : (nseg == 1 ? uq : uq / nseg);
            const long long nb = bundle > 1 ? min((long long)bundle, sp.F - fq) : 1;   // blocks of consecutive frames are contiguous
            const uint32_t nbytes = (uint32_t)nb * (uint32_t)sp.block_bytes;
            mbar_arrive_expect_tx(kfull0 + 8u * slot, nbytes);
            bulk_g2s(ring_s + slot * ring_slot, sp.blocks + (size_t)fq * sp.block_bytes, nbytes, kfull0 + 8u * slot);
            __threadfence_block();
            *issued = t + 1u;
        }
    };

    // per-warp running state: (us, ci) of the next item slot to work on; ring slot / phase of unit us; phase of the row barrier
    uint32_t us = (uint32_t)warp / ips, ci = (uint32_t)warp - us * ips;
    uint32_t kslot = us % ks, kpar = (us / ks) & 1u, rphase = 0u;
    if (us < nu) {
        const Item first = locate(us, ci);
        issue(first, rows_of(first));
        if (lane == 0) ensure_requested((long long)us + look);
    }

    while (us < nu) {
        float x[1][NOFF];
        int st[1], en[1], nrows;
        float score;
        const uint32_t it_us = us, it_ci = ci, it_k = kslot | (kpar << 8);   // all that identifies this item across the evaluation
        {
            const Item cur = locate(us, ci);
            nrows = rows_of(cur);
            if (nrows > 0) {
                mbar_wait(bar_rows, rphase);
                rphase ^= 1u;
            }
            {   // (an item without rows reads whatever the slot holds: nothing of it is used -- its lanes are not `real` -- and an
                // else-branch that zeroes 72 registers costs every item 36 instructions; A/B on one box: 0.829 -> 0.849 of the roofline
                // at the headline shape, 0.75 -> 0.78 at 36 offsets)
                const uintptr_t a0 = (uintptr_t)(sp.props + ((size_t)cur.f * sp.N + cur.r0) * P);
                const uint32_t row = slot_s + 128u + (uint32_t)(a0 & 15) + (uint32_t)(lane < nrows ? lane : 0) * (P * 4);
#pragma unroll
                for (int i = 0; i < NOFF; ++i) x[0][i] = __uint_as_float(lds_u32(row + 4u * (5 + i)));
                st[0] = lane_start(__uint_as_float(lds_u32(row + 8u)), NOFF);            // nms_kernel.cu:29-30
                en[0] = lane_end(__uint_as_float(lds_u32(row + 16u)), st[0], NOFF);      // :32-34
                score = __uint_as_float(lds_u32(slot_s + 4u * (uint32_t)(lane < nrows ? lane : 0)));
            }
            __syncwarp();   // every lane has read its row: the slot is free for the next item
            // advance to this warp's next item slot and request it
            ci += (uint32_t)nwarps;
            if (ci >= ips) {
                uint32_t d = 1u;
                if (ips >= (uint32_t)nwarps) {
                    ci -= ips;
                } else {
                    d = ci / ips;
                    ci -= d * ips;
                }
                us += d;
                kslot += d;
                while (kslot >= ks) {
                    kslot -= ks;
                    kpar ^= 1u;
                }
            }
            if (us < nu) {
                const Item nxt = locate(us, ci);
                issue(nxt, rows_of(nxt));
            }
            // (measured and rejected: an L2 bulk prefetch -- cp.async.bulk.prefetch.L2 -- of the item after next, to keep more bytes
            // in flight than one staging slot per warp admits: no gain at the headline shape, -2 % at top_k 8, -6 % at 36 offsets)
            if (lane == 0) ensure_requested((long long)it_us + look);
            __syncwarp();
        }

        // ---- the frame's kept lanes (block written by the select kernel: {nk, open, n} + slots) -------------------------------
        spin_until_issued(it_us);   // (the request of this unit's block has gone out: the parity below names the right phase)
        mbar_wait(kfull0 + 8u * (it_k & 0xffu), it_k >> 8);
        uint32_t blk_s = ring_s + (it_k & 0xffu) * ring_slot;
        if (bundle > 1) blk_s += (it_ci / (uint32_t)sp.ipf) * (uint32_t)sp.block_bytes;
        const int nk = min((int)lds_u32(blk_s), sp.top_k);
        const int nvf = (int)lds_u32(blk_s + 8);   // proposals in the frame

        const bool real[1] = {lane < nrows};
        uint32_t key = key_desc(real[0] ? score : 0.0f, sp.sort_model == 1);
        // (torch sort model) a frame of <= 32 proposals is ordered by ATen's unstable bitonic network -- the select kernel's
        // rank keys are then the sorted positions, and so must these be; such a frame is exactly rows 0 .. n-1 of one item
        if (sp.sort_model == 0 && nvf <= 32 && nvf >= 2 && nrows > 0) {
            const int nv = nvf;
            float *bit_key = reinterpret_cast<float *>(smem + sp.off_bit + warp * 384);
            int *bit_val = reinterpret_cast<int *>(bit_key + 32), *bit_ok = bit_val + 32;
            bit_ok[lane] = lane < nv;
            bit_key[lane] = lane < nv ? score : 0.0f;
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
            int mypos = 0;
            for (int q = 0; q < 32; ++q)
                if (bit_val[q] == lane && q < nv) mypos = q;
            key = (uint32_t)mypos;
            __syncwarp();
        }
        uint32_t mb[1][MW], par[1] = {0u};
        range_mask<MW>(st[0], en[0], mb[0]);

        if (nrows > 0) {
            // (the item's first row: it.r0 -- recomputed rather than kept in a register across the passes)
            const u64 myK[1] = {real[0] ? (((u64)key << 32) | (uint32_t)(locate(it_us, it_ci).r0 + lane)) : kNone64};
            for (int k0 = 0; k0 < nk; k0 += NKP) {
                const int cnt = min(NKP, nk - k0);
                if (!stream_eval<NOFF, NKP, 1>(blk_s + kBlkHdr + (uint32_t)k0 * SLOT, cnt, real, myK, st, en, mb, x, sp.thr, par, k0)) {
                    // a pair with a negative common start somewhere in the warp: the exact evaluator, one lane at a time
                    FusedParams fp;
                    fp.thr = sp.thr;
                    const Item cur = locate(it_us, it_ci);
                    auto my_hdr = [&](int) { return sp.props + ((size_t)cur.f * sp.N + (uint32_t)(cur.r0 + lane)) * P; };
                    const unsigned char *blk_g = smem + (blk_s - smem_s);
                    for (int k = 0; k < cnt; ++k) {
                        const unsigned char *const h1[1] = {blk_g + kBlkHdr + (size_t)(k0 + k) * SLOT};
                        bool hit[1][1];
                        freg_eval<NOFF, 1, 1>(fp, cur.f, h1, real, myK, st, en, mb, x, my_hdr, par, hit, k0 + k);
                    }
                }
            }
        }
        // ---- outputs, written once ----------------------------------------------------------------------------------------
        {
            const Item cur = locate(it_us, it_ci);
            if (cur.valid) {
                const int i_out = cur.r0 + lane;
                if (i_out < sp.N) {
                    st_global_cs_u64(sp.parent + (size_t)cur.f * sp.N + i_out, (long long)par[0]);
                    if (i_out >= nk) st_global_cs_u64(sp.keep + (size_t)cur.f * sp.N + i_out, 0ll);   // :139-140
                }
                // an open frame with a proposal nobody covers has more lanes to keep than the select kernel looked for
                if (lds_u32(blk_s + 4) != 0u && __any_sync(0xffffffffu, real[0] && par[0] == 0u)) {
                    if (lane == 0 && atomicExch(sp.flags + cur.f, 1) == 0) sp.list[atomicAdd(sp.ctrs, 1u)] = (int)cur.f;
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(kempty0 + 8u * (it_k & 0xffu));
    }
}

}  // namespace phnms

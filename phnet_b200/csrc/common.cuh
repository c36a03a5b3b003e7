// common.cuh -- device helpers shared by the lane-NMS kernels (sm_100a only).
//
// Arithmetic contract (bit-exactness with PHNet libs/ops/csrc/nms_kernel.cu:26-48, `devIoU`):
//   * every fp32 op is an explicit round-to-nearest intrinsic (__fmul_rn/__fadd_rn): never contracted to FMA
//   * (int)(double) is F2I.F64.TRUNC: truncating, saturating, NaN -> INT_MIN on B200 (what the reference compiles to)
//   * int arithmetic wraps; the offset loop counter is an `unsigned char` in the reference (:38)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace phnms {

typedef unsigned long long u64;

constexpr uint32_t kKeyDead = 0xFFFFFFFFu;
constexpr u64 kNone64 = ~0ull;

// ---- ordering keys --------------------------------------------------------------------------------
// ascending-u32 key == descending score (cub radix twiddle, -0.0 folded onto +0.0).
// nan_first: comparator semantics (GTOp<float,true>): every NaN ties for first place.
__device__ __forceinline__ uint32_t key_desc(float s, bool nan_first) {
    uint32_t u = __float_as_uint(s);
    if (nan_first && (s != s)) return 0u;
    if (u == 0x80000000u) u = 0u;
    const uint32_t asc = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ~asc;
}

__device__ __forceinline__ bool gt_nan(float l, float r) {  // GTOp<float, true>
    return ((l != l) && !(r != r)) || (l > r);
}

// ---- per-lane bounds (nms_kernel.cu:29-34) -----------------------------------------------------------
__device__ __forceinline__ int lane_start(float start_y, int n_off) {
    const float m = __fmul_rn(start_y, (float)(n_off - 1));
    return (int)__dadd_rn((double)m, 0.5);
}

// end, already clamped to n_off-1 (min is associative, :34)
__device__ __forceinline__ int lane_end(float length, int start, int n_off) {
    const float t = __fadd_rn(__fadd_rn((float)start, length), -1.0f);
    const float lm1 = __fadd_rn(length, -1.0f);
    double e = __dadd_rn((double)t, 0.5);
    e = __dsub_rn(e, (lm1 < 0.0f) ? 1.0 : 0.0);
    const int end = (int)e;
    return min(end, n_off - 1);
}

// ---- devIoU for 32 pairs at once (nms_kernel.cu:26-48) --------------------------------------------------
// Every lane of the warp holds one column lane `b` (per-lane bounds sb, eb with eb already clamped to n_off-1)
// and all lanes share the row lane `a` (bounds sa, ea).  The offset loop runs over the warp-wide UNION of the
// per-lane [i0, last] ranges so that `a[i]` is a broadcast load and thread-per-row reads of `b` (odd word stride)
// stay bank-conflict free; each lane adds only inside its own range, in ascending offset order, which is the
// reference's summation order (:38-44).  |a-b| == (a<b ? b-a : a-b) bit for bit (also NaN-ness), and adding it
// with one FADD keeps the reference's fp32 rounding sequence.  Must be called by all 32 lanes (inactive: act=false).
//   kVecA: `a` is 16-byte aligned and readable up to the next multiple of 4 words (LDS.128 broadcast)
//   b must be readable up to the next multiple of 4 words past the row (values there are never used).
template <bool kVecA>
__device__ __forceinline__ bool warp_pair_hit(const float *a, const float *b, bool act, int sa, int ea, int sb,
                                              int eb, float thr) {
    const int start = max(sa, sb);  // :31
    const int end = min(ea, eb);    // :34
    act = act && (end >= start);    // :36
    // :38  for (unsigned char i = 5 + start; i <= 5 + end; ++i)   (counter wraps mod 256 at initialisation only)
    const int i0 = (int)(((uint32_t)start + 5u) & 255u);
    const int last = (int)((uint32_t)end + 5u);
    const bool run = act && (i0 <= last);
    const int lo = run ? i0 : (1 << 20);
    const int hi = run ? last : -1;
    const int wmin = __reduce_min_sync(0xffffffffu, lo);
    const int wmax = __reduce_max_sync(0xffffffffu, hi);
    float dist = 0.0f;
    for (int i4 = wmin & ~3; i4 <= wmax; i4 += 4) {
        float a0, a1, a2, a3;
        if (kVecA) {
            const float4 av = *reinterpret_cast<const float4 *>(a + i4);
            a0 = av.x; a1 = av.y; a2 = av.z; a3 = av.w;
        } else {
            a0 = a[i4]; a1 = a[i4 + 1]; a2 = a[i4 + 2]; a3 = a[i4 + 3];
        }
        const float b0 = b[i4], b1 = b[i4 + 1], b2 = b[i4 + 2], b3 = b[i4 + 3];
        if (i4 >= lo && i4 <= hi) dist = __fadd_rn(dist, fabsf(__fsub_rn(a0, b0)));
        if (i4 + 1 >= lo && i4 + 1 <= hi) dist = __fadd_rn(dist, fabsf(__fsub_rn(a1, b1)));
        if (i4 + 2 >= lo && i4 + 2 <= hi) dist = __fadd_rn(dist, fabsf(__fsub_rn(a2, b2)));
        if (i4 + 3 >= lo && i4 + 3 <= hi) dist = __fadd_rn(dist, fabsf(__fsub_rn(a3, b3)));
    }
    const int len = (int)((uint32_t)end - (uint32_t)start + 1u);
    return act && (dist < __fmul_rn(thr, (float)len));  // :46
}

// One pair, one thread (used where only a handful of pairs is needed, e.g. candidate-vs-candidate in topm.cuh).
// sb/eb and sa/ea are the lanes' own bounds with the ends already clamped to n_off-1.
__device__ __forceinline__ bool pair_hit_scalar(const float *a, const float *b, int sa, int ea, int sb, int eb, float thr) {
    const int start = max(sa, sb), end = min(ea, eb);              // nms_kernel.cu:31,34
    if (end < start) return false;                                  // :36
    const int i0 = (int)(((uint32_t)start + 5u) & 255u);            // :38 unsigned char counter
    const int last = (int)((uint32_t)end + 5u);
    const int len = (int)((uint32_t)end - (uint32_t)start + 1u);
    const float lim = __fmul_rn(thr, (float)len);                   // :46
    float dist = 0.0f;
    int i = i0;
    // Exact early exit: every term is >= 0 and fp32 addition is monotone, so once the running sum has reached the limit the
    // final `dist < lim` is false whatever follows (a NaN makes both comparisons false and the loop simply runs on).
    for (; i + 3 <= last; i += 4) {   // loads of four offsets in flight, the adds stay sequential and in order
        const float a0 = a[i], a1 = a[i + 1], a2 = a[i + 2], a3 = a[i + 3];
        const float b0 = b[i], b1 = b[i + 1], b2 = b[i + 2], b3 = b[i + 3];
        dist = __fadd_rn(dist, fabsf(__fsub_rn(a0, b0)));
        dist = __fadd_rn(dist, fabsf(__fsub_rn(a1, b1)));
        dist = __fadd_rn(dist, fabsf(__fsub_rn(a2, b2)));
        dist = __fadd_rn(dist, fabsf(__fsub_rn(a3, b3)));
        if (dist >= lim) return false;
    }
    for (; i <= last; ++i) dist = __fadd_rn(dist, fabsf(__fsub_rn(a[i], b[i])));
    return dist < lim;
}

// ---- warp helpers --------------------------------------------------------------------------------
__device__ __forceinline__ u64 warp_min_u64(u64 v) {
    const uint32_t hi = (uint32_t)(v >> 32), lo = (uint32_t)v;
    const uint32_t mhi = __reduce_min_sync(0xffffffffu, hi);
    const uint32_t mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
    return ((u64)mhi << 32) | mlo;
}

// ---- shared-memory / cluster / mbarrier / bulk-copy PTX ----------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_arrive_release() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_acquire() {
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded: a wait that has not completed after kMbarTimeoutNs traps (the launch fails with an error the caller sees)
// instead of hanging the GPU -- a protocol bug must never become a hung box.  The clock is only read on the slow path.
constexpr unsigned long long kMbarTimeoutNs = 4000000000ull;
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0u;
}
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
    unsigned long long t0, now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        for (int i = 0; i < 64; ++i)
            if (mbar_try_wait(bar, parity)) return;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (now - t0 > kMbarTimeoutNs) {
#ifdef PHNMS_WAIT_PRINT   // debugging build: say who waited on what and carry on (results are wrong, the launch terminates)
            printf("phnms: mbarrier wait timed out: block %d thread %d barrier@%u parity %u\n", (int)blockIdx.x, (int)threadIdx.x,
                   bar, parity);
            return;
#else
            __trap();
#endif
        }
    }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;   // (try_wait itself suspends the warp for a while before it gives up)
    if (mbar_try_wait(bar, parity)) return;
    mbar_wait_slow(bar, parity);
}
// Debug variant: gives up after ~2^22 polls and appends {tag, block, thread, parity, a, b} to `dbg` (first word = count).
__device__ __forceinline__ bool mbar_wait_watch(uint32_t bar, uint32_t parity, long long *dbg, int tag, long long a,
                                                long long b) {
    for (unsigned it = 0; it < (1u << 22); ++it) {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return true;
    }
    const unsigned long long slot = atomicAdd(reinterpret_cast<unsigned long long *>(dbg), 1ull);
    if (slot < 200) {
        long long *r = dbg + 8 + slot * 8;
        r[0] = tag; r[1] = blockIdx.x; r[2] = threadIdx.x; r[3] = parity; r[4] = a; r[5] = b;
    }
    return false;
}
// TMA 1-D bulk copy global -> shared (UBLKCP); completes `bytes` on the mbarrier.
// src and dst 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// Compact kept-lane records {keep[0 .. top_k), num_to_keep} of a frame, stored by the NMS kernels themselves into every listed
// destination (local memory or other GPUs' buffers, then over NVLink): n == 0 switches it off.  See phnms_forward_collect_f32.
constexpr int kMaxCollectDst = 16;   // == PHNMS_MAX_DST
struct RecordSink {
    int n, width;        // destinations; int64 words per record (top_k + 1)
    long long row0;      // this call's frame f goes to row row0 + f of every destination
    long long *dst[kMaxCollectDst];
};
__device__ __forceinline__ void record_store(const RecordSink &r, long long f, int c, long long v) {
    for (int d = 0; d < r.n; ++d) r.dst[d][(r.row0 + f) * r.width + c] = v;
}

__device__ __forceinline__ void st_global_cs_u64(long long *p, long long v) {  // streaming store: outputs are write-once
    asm volatile("st.global.cs.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

}  // namespace phnms

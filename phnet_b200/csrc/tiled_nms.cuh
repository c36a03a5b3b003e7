// tiled_nms.cuh -- the general path: three kernels, any N the reference accepts (ceil(N/64) < 1000).
//
//   (1) order_kernel   per-frame descending order of the scores            <- libs/ops/csrc/nms.cpp:51  scores.sort(0,true)
//                      warp-level LSD radix sort: 8-bit digits, per-warp histograms, __match_any_sync multisplit
//                      (stable, so equal scores keep ascending index); n <= 32 replays ATen's bitonic network.
//   (2) mask_kernel    64x64 tile of the pair predicate -> one u64 per (row, column block), upper triangle only
//                      <- nms_kernel.cu:50-96.  Rows are gathered through `order` with 128-bit global loads of the
//                      16-byte aligned window around each (4-byte aligned) row and staged in shared memory.
//   (3) scan_kernel    one warp per frame: greedy suppression over the bitmask with ballot/ffs
//                      <- nms_kernel.cu:99-143 (nms_collect, a single thread in the reference).
//
// This path materialises the N x ceil(N/64) bitmask in HBM and evaluates all N(N-1)/2 pairs, so it is bound by
// fp32 issue rate, not by HBM.  The fused path (fused_nms.cuh) is used whenever a frame fits a cluster.
#pragma once
#include "common.cuh"

namespace phnms {

// ------------------------------------------------------------------------------------------------------
// (1) ordering
// ------------------------------------------------------------------------------------------------------
constexpr int kOrderThreads = 512;
constexpr int kOrderWarps = kOrderThreads / 32;

// ws per frame: 2 key buffers + 2 index buffers of N u32
__global__ void __launch_bounds__(kOrderThreads) phnms_order_kernel(const float *__restrict__ scores,
                                                                   const int32_t *__restrict__ n_valid, int N,
                                                                   int sort_model, long long *__restrict__ order,
                                                                   uint32_t *__restrict__ ws) {
    __shared__ uint32_t hist[kOrderWarps][256];
    __shared__ uint32_t dig_total[256];
    __shared__ float bit_key[32];
    __shared__ int bit_val[32];
    __shared__ int bit_ok[32];

    const long long f = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int n = N;
    if (n_valid) n = max(0, min(n_valid[f], N));
    const float *sc = scores + (size_t)f * N;
    long long *out = order + (size_t)f * N;

    if (n <= 1 || (sort_model == 0 && n <= 32)) {
        if (warp == 0) {
            bit_ok[lane] = lane < n;
            bit_key[lane] = lane < n ? sc[lane] : 0.0f;
            bit_val[lane] = lane < n ? lane : 0;
            __syncwarp();
            if (n > 1) {  // ATen bitonicSortKVInPlace<block_dim_x = 16> (SortUtils.cuh:45-163)
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
            }
        }
        __syncthreads();
        for (int i = tid; i < N; i += kOrderThreads) out[i] = (i < n && i < 32) ? (long long)bit_val[i] : 0ll;
        return;
    }

    const bool nan_first = sort_model == 1;
    uint32_t *k0 = ws + (size_t)f * 4 * N, *k1 = k0 + N, *v0 = k1 + N, *v1 = v0 + N;
    for (int i = tid; i < n; i += kOrderThreads) {
        k0[i] = key_desc(sc[i], nan_first);
        v0[i] = (uint32_t)i;
    }
    __syncthreads();

    // each warp owns a contiguous segment (multiple of 32 long) so that the scatter order is the input order
    const int seg = (((n + kOrderWarps - 1) / kOrderWarps) + 31) & ~31;
    const int s0 = min(warp * seg, n), s1 = min(s0 + seg, n);

    for (int pass = 0; pass < 4; ++pass) {
        const int shift = pass * 8;
        for (int i = tid; i < kOrderWarps * 256; i += kOrderThreads) (&hist[0][0])[i] = 0u;
        __syncthreads();
        for (int i = s0 + lane; i < s1; i += 32) atomicAdd(&hist[warp][(k0[i] >> shift) & 255u], 1u);
        __syncthreads();
        // exclusive scan in (digit major, warp minor) order
        if (tid < 256) {
            uint32_t sum = 0;
            for (int w = 0; w < kOrderWarps; ++w) {
                const uint32_t c = hist[w][tid];
                hist[w][tid] = sum;
                sum += c;
            }
            dig_total[tid] = sum;
        }
        __syncthreads();
        if (warp == 0) {
            uint32_t carry = 0;
            for (int d0 = 0; d0 < 256; d0 += 32) {
                const uint32_t v = dig_total[d0 + lane];
                uint32_t inc = v;
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += t;
                }
                dig_total[d0 + lane] = carry + inc - v;
                carry += __shfl_sync(0xffffffffu, inc, 31);
            }
        }
        __syncthreads();
        // stable scatter: groups of 32 in input order, rank inside the group by __match_any_sync
        for (int i = s0; i < s1; i += 32) {
            const int j = i + lane;
            const bool ok = j < s1;
            const uint32_t key = ok ? k0[j] : 0u, val = ok ? v0[j] : 0u;
            const uint32_t d = ok ? ((key >> shift) & 255u) : 256u + lane;  // inactive lanes match nobody
            const uint32_t peers = __match_any_sync(0xffffffffu, d);
            const uint32_t before = __popc(peers & ((1u << lane) - 1u));
            uint32_t base = 0;
            if (ok) base = dig_total[d] + hist[warp][d];
            __syncwarp();
            if (ok) {
                k1[base + before] = key;
                v1[base + before] = val;
                if (before == 0) hist[warp][d] += __popc(peers);
            }
            __syncwarp();
        }
        __syncthreads();
        uint32_t *t = k0; k0 = k1; k1 = t;
        t = v0; v0 = v1; v1 = t;
    }
    for (int i = tid; i < N; i += kOrderThreads) out[i] = i < n ? (long long)v0[i] : 0ll;
}

// ------------------------------------------------------------------------------------------------------
// (2) pair-predicate bitmask, 64 x 64 tiles
// ------------------------------------------------------------------------------------------------------
constexpr int kTile = 64;
constexpr int kMaskThreads = 256;

// dynamic smem: 128 rows x stride words + 128 x int2 bounds; stride = round4(P) + 1 (odd -> conflict free)
__global__ void __launch_bounds__(kMaskThreads) phnms_mask_kernel(const float *__restrict__ props,
                                                                 const long long *__restrict__ order,
                                                                 const int32_t *__restrict__ n_valid, int N, int n_off,
                                                                 float thr, int col_blocks,
                                                                 unsigned long long *__restrict__ mask,
                                                                 const float *__restrict__ props_end) {
    extern __shared__ __align__(128) unsigned char smem[];
    const long long f = blockIdx.y;
    int n = N;
    if (n_valid) n = max(0, min(n_valid[f], N));
    // linear tile id -> (row block rb <= column block cb) of the upper triangle
    int rb = 0, rem = blockIdx.x;
    while (rem >= col_blocks - rb) { rem -= col_blocks - rb; ++rb; }
    const int cb = rb + rem;
    if (rb * kTile >= n || cb * kTile >= n) return;

    const int P = 5 + n_off;
    const int stride = ((P + 3) & ~3) + 1;
    float *tile = reinterpret_cast<float *>(smem);                      // [128][stride]: 0..63 columns, 64..127 rows
    int2 *bounds = reinterpret_cast<int2 *>(smem + (size_t)2 * kTile * stride * 4);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int col_size = min(n - cb * kTile, kTile), row_size = min(n - rb * kTile, kTile);
    const long long *ord = order + (size_t)f * N;
    const float *frame = props + (size_t)f * N * P;

    // stage 128 proposals: one warp per proposal, 128-bit loads over the aligned window that covers the row
    for (int r = warp; r < 2 * kTile; r += kMaskThreads / 32) {
        const bool is_row = r >= kTile;
        const int local = is_row ? r - kTile : r;
        const int sorted = (is_row ? rb : cb) * kTile + local;
        float *dst = tile + (size_t)r * stride;
        if (local < (is_row ? row_size : col_size)) {
            const float *src = frame + (size_t)ord[sorted] * P;
            const uintptr_t s = (uintptr_t)src, s_al = s & ~(uintptr_t)15;
            const int head = (int)((s - s_al) >> 2);  // words of the previous row in front of this one
            const int nvec = (head + P + 3) >> 2;
            for (int v = lane; v < nvec; v += 32) {
                const float *g = reinterpret_cast<const float *>(s_al) + 4 * v;
                float4 q;
                if (g + 4 <= props_end) {
                    q = __ldg(reinterpret_cast<const float4 *>(g));
                } else {  // last row of the tensor: never read past the allocation
                    q.x = g + 0 < props_end ? g[0] : 0.f;
                    q.y = g + 1 < props_end ? g[1] : 0.f;
                    q.z = g + 2 < props_end ? g[2] : 0.f;
                    q.w = g + 3 < props_end ? g[3] : 0.f;
                }
                const int w = 4 * v - head;
                if (w >= 0 && w < P) dst[w] = q.x;
                if (w + 1 >= 0 && w + 1 < P) dst[w + 1] = q.y;
                if (w + 2 >= 0 && w + 2 < P) dst[w + 2] = q.z;
                if (w + 3 >= 0 && w + 3 < P) dst[w + 3] = q.w;
            }
            __syncwarp();
            if (lane == 0) {
                const int st = lane_start(dst[2], n_off);
                bounds[r] = make_int2(st, lane_end(dst[4], st, n_off));
            }
        } else if (lane == 0) {
            bounds[r] = make_int2(0, -1);
        }
    }
    __syncthreads();

    // warp w evaluates tile rows 8w .. 8w+7 against all 64 columns (two per lane), ballot -> one u64 per row
    for (int rr = 0; rr < kTile / (kMaskThreads / 32); ++rr) {
        const int r = warp * (kTile / (kMaskThreads / 32)) + rr;
        if (r >= row_size) break;
        const float *a = tile + (size_t)(kTile + r) * stride;
        const int2 sea = bounds[kTile + r];
        const int first = (rb == cb) ? r + 1 : 0;  // strict upper triangle (nms_kernel.cu:85-87)
        uint32_t w0, w1;
        {
            const int c = lane;
            const bool act = c >= first && c < col_size;
            const int2 seb = bounds[c];
            w0 = __ballot_sync(0xffffffffu, warp_pair_hit<false>(a, tile + (size_t)c * stride, act, sea.x, sea.y, seb.x, seb.y, thr));
        }
        {
            const int c = lane + 32;
            const bool act = c >= first && c < col_size;
            const int2 seb = bounds[c];
            w1 = __ballot_sync(0xffffffffu, warp_pair_hit<false>(a, tile + (size_t)c * stride, act, sea.x, sea.y, seb.x, seb.y, thr));
        }
        if (lane == 0)
            mask[((size_t)f * N + (size_t)rb * kTile + r) * col_blocks + cb] = ((unsigned long long)w1 << 32) | w0;
    }
}

// ------------------------------------------------------------------------------------------------------
// (3) greedy scan, one warp per frame
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) phnms_scan_kernel(const long long *__restrict__ order,
                                                       const unsigned long long *__restrict__ mask,
                                                       const int32_t *__restrict__ n_valid, int N, int col_blocks,
                                                       long long top_k, long long *__restrict__ keep,
                                                       long long *__restrict__ num_keep,
                                                       long long *__restrict__ parent) {
    __shared__ unsigned long long remv[1000];  // MAX_COL_BLOCKS (nms_kernel.cu:10,100)
    const long long f = blockIdx.x;
    const int lane = threadIdx.x;
    int n = N;
    if (n_valid) n = max(0, min(n_valid[f], N));
    const long long *ord = order + (size_t)f * N;
    long long *keep_f = keep + (size_t)f * N, *par_f = parent + (size_t)f * N;
    const int cbn = (n + 63) >> 6;
    for (int w = lane; w < cbn; w += 32) remv[w] = 0ull;        // :103-105
    for (int i = lane; i < N; i += 32) par_f[i] = 0ll;          // :107-109
    __syncwarp();
    __threadfence_block();

    long long nk = 0;
    int i = 0;  // next sorted position to examine
    while (i < n) {
        // first position >= i whose removed bit is clear (:116)
        int found = -1;
        for (int wb = i >> 6; wb < cbn && found < 0; wb += 32) {
            const int w = wb + lane;
            unsigned long long free_bits = 0ull;
            if (w < cbn) {
                free_bits = ~remv[w];
                if (w == (i >> 6)) free_bits &= ~0ull << (i & 63);
                const int top = n - (w << 6);
                if (top < 64) free_bits &= (1ull << top) - 1ull;
            }
            const uint32_t any = __ballot_sync(0xffffffffu, free_bits != 0ull);
            if (any) {
                const int src = __ffs(any) - 1;
                const int pos = ((wb + src) << 6) + (__ffsll((long long)__shfl_sync(0xffffffffu, free_bits, src)) - 1);
                found = pos;
            }
        }
        if (found < 0) break;
        i = found;
        const int nblock = i >> 6;
        const long long idxi = ord[i];
        const unsigned long long *row = mask + ((size_t)f * N + i) * col_blocks;
        for (int w = nblock + lane; w < cbn; w += 32) {
            unsigned long long m = row[w];
            remv[w] |= m;                                       // :120-122
            while (m) {                                         // :123-128
                const int b = __ffsll((long long)m) - 1;
                m &= m - 1;
                par_f[ord[(w << 6) + b]] = nk + 1;
            }
        }
        if (lane == 0) {
            keep_f[nk] = idxi;                                  // :118
            par_f[idxi] = nk + 1;                               // :129
        }
        ++nk;
        ++i;
        __syncwarp();
        if (nk == top_k) break;                                 // :133
    }
    __syncwarp();
    for (long long j = nk + lane; j < N; j += 32) keep_f[j] = 0ll;  // :139-140
    if (lane == 0) num_keep[f] = top_k < nk ? top_k : nk;            // :142
}

// ---- double precision boxes (the reference also instantiates nms_kernel<double>, nms_kernel.cu:171) --------------------
// A compatibility path, not a fast one: one thread per (sorted row, column block) evaluates its 64 pairs with scalar
// loops straight from global memory; phnms_scan_kernel then runs unchanged on the bitmask.  The arithmetic follows the
// SASS of the reference's double instantiation (oracle/_ref): the start is ONE fused multiply-add
// (`a[2] * N_STRIPS + 0.5` contracts to DFMA under nvcc's default -fmad=true), the end is a chain of DADDs, both through
// F2I.F64.TRUNC; the distance is a sequential fp64 sum; the limit is the FP32 product thr * len widened to double
// (devIoU takes `const float threshold`, :26).
__device__ __forceinline__ int lane_start_f64(double y, int n_off) {
    return __double2int_rz(__fma_rn(y, (double)(n_off - 1), 0.5));                          // :29-30
}
__device__ __forceinline__ int lane_end_f64(double len, int start) {                         // :32-33
    double e = __dadd_rn(__dadd_rn(__dadd_rn((double)start, len), -1.0), 0.5);
    e = __dsub_rn(e, (__dadd_rn(len, -1.0) < 0.0) ? 1.0 : 0.0);
    return __double2int_rz(e);
}
__device__ __forceinline__ bool pair_hit_f64(const double *a, const double *b, int sa, int ea, int sb, int eb, int n_off,
                                             float thr) {
    const int start = max(sa, sb);                                                           // :31
    const int end = min(min(ea, eb), n_off - 1);                                             // :34
    if (end < start) return false;                                                           // :36
    const int i0 = (int)(((uint32_t)start + 5u) & 255u);                                     // :38 unsigned char counter
    const int last = (int)((uint32_t)end + 5u);
    double dist = 0.0;
    for (int i = i0; i <= last; ++i) {
        const double av = a[i], bv = b[i];
        dist = __dadd_rn(dist, (av < bv) ? __dsub_rn(bv, av) : __dsub_rn(av, bv));           // :39-43
    }
    const int len = (int)((uint32_t)end - (uint32_t)start + 1u);
    return dist < (double)__fmul_rn(thr, (float)len);                                        // :46
}

__global__ void __launch_bounds__(128) phnms_mask_f64_kernel(const double *__restrict__ props,
                                                            const long long *__restrict__ order,
                                                            const int32_t *__restrict__ n_valid, int N, int n_off,
                                                            float thr, int col_blocks,
                                                            unsigned long long *__restrict__ mask) {
    const long long f = blockIdx.z;
    const int w = blockIdx.y, i = blockIdx.x * 128 + threadIdx.x, P = 5 + n_off;
    int n = N;
    if (n_valid) n = max(0, min(n_valid[f], N));
    if (i >= n || w < (i >> 6)) return;                                                      // :56 upper triangle only
    const long long *ord = order + (size_t)f * N;
    const double *a = props + ((size_t)f * N + ord[i]) * P;
    const int sa = lane_start_f64(a[2], n_off), ea = lane_end_f64(a[4], sa);
    unsigned long long bits = 0ull;
    for (int b = 0; b < 64; ++b) {
        const int j = 64 * w + b;
        if (j <= i || j >= n) continue;                                                      // :85-87 strict upper triangle
        const double *q = props + ((size_t)f * N + ord[j]) * P;
        const int sb = lane_start_f64(q[2], n_off), eb = lane_end_f64(q[4], sb);
        if (pair_hit_f64(a, q, sa, ea, sb, eb, n_off, thr)) bits |= 1ull << b;
    }
    mask[((size_t)f * N + i) * col_blocks + w] = bits;                                       // :94
}

}  // namespace phnms

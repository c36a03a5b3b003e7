// phnms.cu -- C ABI (include/phnms.h) and launch planning for the B200 lane-NMS library.
//
// Built with:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared -Xcompiler -fPIC
// No torch, no ATen: the Python mirror (phnet_b200/ops/nms.py) passes raw device pointers and a stream.
#include "../../include/phnms.h"

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "fused_nms.cuh"
#include "fused_reg.cuh"
#include "tiled_nms.cuh"
#include "topm.cuh"
#include "select.cuh"
#include "stream.cuh"
#include "frontend.cuh"
#include "small.cuh"
#include "assign.cuh"

using namespace phnms;

namespace {

struct DeviceInfo {
    int sms;
    int smem_optin;  // max dynamic shared memory per CTA
    int cc_major;
};

// Device attributes are cached per device: a per-frame call (the way PHNet calls the op) must not pay three attribute
// queries each time.  Benign race: concurrent first calls write identical values.
int device_info(DeviceInfo *d) {
    static DeviceInfo cache[64];
    static volatile int ready[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev >= 0 && dev < 64 && ready[dev]) {
        *d = cache[dev];
        return 0;
    }
    cudaDeviceGetAttribute(&d->sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&d->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    e = cudaDeviceGetAttribute(&d->cc_major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e == cudaSuccess && dev >= 0 && dev < 64) {
        cache[dev] = *d;
        __sync_synchronize();
        ready[dev] = 1;
    }
    return (int)e;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device): raised to the device's opt-in maximum.
// (keyed by the function's address: several template instances share one pointer TYPE)
int ensure_max_smem(const void *kern, int smem_optin) {
    struct Entry { const void *fn; int dev; };
    static Entry table[256];
    static int count = 0;
    static std::mutex mu;
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> g(mu);
        for (int i = 0; i < count; ++i)
            if (table[i].fn == kern && table[i].dev == dev) return 0;
    }
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, kern);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin - (int)fa.sharedSizeBytes);
    if (e != cudaSuccess) return (int)e;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) cudaGetLastError();
    std::lock_guard<std::mutex> g(mu);
    if (count < 256) table[count++] = Entry{kern, dev};
    return 0;
}

// experiment knobs of debug sessions, read once at load (never on the call path)
struct EnvKnobs {
    int topm_count, no_topm, select_cap, lanes_per_pass, stream_warps, stream_cpt, no_stream, no_small, small_grid_f, no_fused_records, debug, skip, small_batch;
    EnvKnobs() {
        auto geti = [](const char *n) { const char *v = getenv(n); return v ? atoi(v) : 0; };
        topm_count = geti("PHNMS_TOPM_COUNT");
        no_topm = getenv("PHNMS_NO_TOPM") != nullptr;
        select_cap = geti("PHNMS_SELECT_CAP");
        lanes_per_pass = geti("PHNMS_LANES_PER_PASS");
        stream_warps = geti("PHNMS_STREAM_WARPS");
        stream_cpt = geti("PHNMS_STREAM_CPT");
        no_stream = getenv("PHNMS_NO_STREAM") != nullptr;
        no_small = getenv("PHNMS_NO_SMALL") != nullptr;
        small_grid_f = getenv("PHNMS_SMALL_GRID_F") != nullptr;
        no_fused_records = getenv("PHNMS_NO_FUSED_RECORDS") != nullptr;   // A/B: records by the separate collect kernel   // experiment: one CTA per frame instead of persistent CTAs
        debug = getenv("PHNMS_DEBUG") != nullptr;
        small_batch = getenv("PHNMS_SMALL_BATCH") ? geti("PHNMS_SMALL_BATCH") : -1;
        skip = geti("PHNMS_SKIP");   // timing experiments only (results are wrong): 1 = no select, 2 = no stream, 4 = no resume
    }
};
const EnvKnobs g_env;

// PHNMS_DEBUG=1: say on stderr which step of a call failed (the return code alone does not)
int fail_at(const char *step, int code) {
    if (code != 0 && g_env.debug) fprintf(stderr, "phnms: %s failed: %d (%s)\n", step, code, code > 0 ? cudaGetErrorString((cudaError_t)code) : "argument");
    return code;
}

int check_shape(int64_t F, int64_t N, int n_off) {
    if (F < 0 || N < 0) return PHNMS_ERR_BAD_ARG;
    if (n_off < 1 || n_off > 250) return PHNMS_ERR_N_OFFSETS;
    if ((N + 63) / 64 >= 1000) return PHNMS_ERR_TOO_MANY;  // nms_kernel.cu:158
    return PHNMS_OK;
}

size_t tiled_workspace(int64_t F, int64_t N) {
    const size_t cb = (size_t)((N + 63) / 64);
    // order (i64) + radix ping-pong (4 x u32) + bitmask
    return (size_t)F * N * 8 + (size_t)F * N * 16 + (size_t)F * N * cb * 8 + 256;
}

constexpr long long kSmallBatchProposals = 2048;
constexpr int kSmallAnyBatchN = 256;   // frames up to this size take the one-launch kernel (small.cuh) whatever the batch size
constexpr int kSelFullScanN = 256;   // F * N at or below this: single-launch path (see make_plan)

// Launch shape of the streaming kernel (stream.cuh): warps per CTA, how frames are cut into units when there are fewer
// frames than SMs, the kept-block ring, the grid (one persistent CTA per SM).
struct StreamShape {
    int warps, cpt, lanes, ipf, nseg, ips, bundle, ks, block_bytes, grid;
    StreamLayout L;
    bool ok, bad_tuning;
};

StreamShape stream_shape(int64_t F, int64_t N, int n_off, int64_t top_k, const phnms_tuning &t, int sms, int smem_max) {
    StreamShape ss = {};
    const int P = 5 + n_off;
    ss.warps = t.stream_warps ? t.stream_warps : (g_env.stream_warps ? g_env.stream_warps : kStreamMaxWarps);
    if (ss.warps < 1 || ss.warps > kStreamMaxWarps) {
        ss.bad_tuning = true;
        return ss;
    }
    // rows per thread: one.  (Two rows per thread at n_off 36 -- 72 offset registers either way, items twice as large -- was
    // measured slower: 0.64 vs 0.73 of the roofline at N = 1000, top_k = 4.)
    ss.cpt = 1;
    ss.lanes = t.lanes_per_pass ? t.lanes_per_pass : g_env.lanes_per_pass;
    // measured (N = 1000): 72 offsets: 2 lanes per pass 0.86 of the roofline, 4 lanes 0.78 (spills); 36 offsets: 4 lanes 0.775, 2 lanes 0.76;
    // three lanes per pass (top_k 5..8) spill as well: 0.46 vs 0.54 at top_k 8, 0.65 vs 0.71 at top_k 5
    if (ss.lanes == 0) ss.lanes = (top_k >= 0 && top_k < 2) ? 1 : ((n_off == 36 && (top_k < 0 || top_k >= 4)) ? 4 : 2);
    if (ss.lanes != 1 && ss.lanes != 2 && ss.lanes != 4) {
        ss.bad_tuning = true;
        return ss;
    }
    ss.ipf = (int)((N + 32 * ss.cpt - 1) / (32 * ss.cpt));
    if (ss.ipf < 1) ss.ipf = 1;   // (N == 0 is answered before any launch; keep the arithmetic below defined)
    const int kk = top_k < 0 ? kStreamMaxK : (int)top_k;
    ss.block_bytes = kBlkHdr + kk * (kHdr + 4 * ((P + 3) & ~3));
    ss.nseg = 1;
    ss.bundle = 1;
    if (F < sms) {   // fewer frames than SMs: cut a frame into segments so that every SM has work
        ss.nseg = (int)((sms + F - 1) / (F > 0 ? F : 1));
        if (ss.nseg > ss.ipf) ss.nseg = ss.ipf;
        if (ss.nseg < 1) ss.nseg = 1;
    } else if (ss.ipf < ss.warps) {
        // small frames: a unit is a bundle of consecutive frames, so that every warp of the CTA has an item in every unit (the
        // warps then move through the kept-block ring together); bounded by the shared memory the ring may take
        ss.bundle = (ss.warps + ss.ipf - 1) / ss.ipf;
        const int max_bundle = (24 * 1024) / (4 * ss.block_bytes);
        if (ss.bundle > max_bundle) ss.bundle = max_bundle;
        if ((long long)ss.bundle * sms > F) ss.bundle = (int)(F / sms);   // (never fewer units than SMs)
        if (ss.bundle < 1) ss.bundle = 1;
    }
    ss.ips = ss.bundle > 1 ? ss.bundle * ss.ipf : (ss.ipf + ss.nseg - 1) / ss.nseg;
    if (ss.bundle == 1) ss.nseg = (ss.ipf + ss.ips - 1) / ss.ips;
    ss.ks = (ss.warps + ss.ips - 1) / ss.ips + 3;
    if (ss.ks < 4) ss.ks = 4;
    if (ss.ks > kStreamMaxRing) ss.ks = kStreamMaxRing;
    ss.L = stream_layout(ss.warps, P, ss.block_bytes, ss.ks, ss.cpt, ss.bundle);
    while (ss.L.total > smem_max && ss.ks > 4) ss.L = stream_layout(ss.warps, P, ss.block_bytes, --ss.ks, ss.cpt, ss.bundle);
    while (ss.L.total > smem_max && ss.warps > 1) ss.L = stream_layout(--ss.warps, P, ss.block_bytes, ss.ks, ss.cpt, ss.bundle);
    const long long units = ss.bundle > 1 ? (F + ss.bundle - 1) / ss.bundle : (long long)F * ss.nseg;
    long long grid = units < sms ? units : sms;
    if (t.max_clusters > 0 && grid > t.max_clusters) grid = t.max_clusters;
    if (grid < 1) grid = 1;
    ss.grid = (int)grid;
    ss.ok = ss.L.total <= smem_max && ((units + grid - 1) / grid + 1) * ss.ips < 0x7fffffffLL;
    return ss;
}

// Decide what to launch.  No device queries when `dev` is null (shape-only planning with B200 constants).
// top_k < 0: unknown (shape-only queries) -- assume PHNet's range [1, 8].
int make_plan(int64_t F, int64_t N, int n_off, int64_t top_k, const phnms_tuning *tun, const DeviceInfo *dev, phnms_plan *pl) {
    const int P = 5 + n_off;
    const int smem_max = dev ? dev->smem_optin : 232448;
    const int sms = dev ? dev->sms : 148;
    phnms_tuning t = phnms_tuning{};
    if (tun) t = *tun;
    pl->workspace_bytes = 0;
    pl->launches = 1;
    pl->variant = 0;
    pl->cols_per_thread = 1;
    pl->max_active_clusters = 0;
    const int cand[5] = {1, 2, 4, 8, 16};

    auto rows_for = [&](int c) {
        int r = (int)((N + c - 1) / c);
        return r < 32 ? 32 : r;  // a frame of <= 32 proposals lives in one CTA (bitonic replay needs all of them)
    };

    if (t.path != PHNMS_PATH_TILED) {
        // ---- register-resident variant: n_off 36 / 72 only; a CTA holds at most 512 threads x cols_per_thread rows
        const bool reg_ok = (n_off == 36 || n_off == 72) && t.variant != PHNMS_FUSED_SMEM;
        // The streaming path: the default when nothing asks for the cluster kernels.  Its resume pass IS the register-resident
        // cluster kernel, so it needs that plan to exist (N <= 8192).
        const bool k_ok = top_k < 0 || (top_k >= 1 && top_k <= kStreamMaxK);
        // A call with only a handful of proposals (PHNet's own call: ONE frame of <= 240) is launch-latency bound: it takes
        // the register-resident cluster kernel with in-kernel selection -- one launch instead of three.
        const long long small_limit = g_env.small_batch >= 0 ? g_env.small_batch : kSmallBatchProposals;
        const bool small = (long long)F * N <= small_limit && t.variant == 0;
        const bool want_stream = t.variant == PHNMS_FUSED_STREAM ||
                                 (t.variant == 0 && !t.cluster && !t.threads && !t.schedule && !g_env.no_stream && !small);
        if (t.variant == PHNMS_FUSED_STREAM && (!reg_ok || !k_ok || t.cluster || t.threads || t.schedule)) return PHNMS_ERR_TUNING;
        // ... and when its frames have at most 512 proposals, the one-launch kernel of small.cuh: one CTA per frame, no workspace.
        // Frames of at most 256 proposals (everything PHNet itself produces: 240 priors) take it whatever the batch size: persistent
        // CTAs, the next frame's rows in flight during the greedy rounds of the current one; its cost does not depend on where in
        // the order the kept lanes sit, and any top_k goes (measured against select -> stream at N = 240: 0.68 vs 0.56-0.66 of the
        // roofline at 72 offsets / top_k 4, 0.32 vs 0.22-0.27 at 36 offsets / top_k 8, before the prefetch).
        const bool small_ok = reg_ok && N <= kSmallMaxN && !t.cluster && !t.threads && !t.schedule;
        if (t.variant == PHNMS_FUSED_SMALL && !small_ok) return PHNMS_ERR_TUNING;
        if (small_ok && (t.variant == PHNMS_FUSED_SMALL || (t.variant == 0 && !g_env.no_small && (small || N <= kSmallAnyBatchN))) &&
            F <= 0x7fffffff) {
            const int warps = N > 32 ? (int)((N + 31) / 32) : 1;
            const int smem = small_layout(warps, P).total;
            // resident CTAs per SM: shared memory, registers (128 per thread; 80 for <= 256 threads at 36 offsets), threads
            const int regs = (n_off == 36 && warps <= 8) ? 80 : 128;
            int per_sm = (smem_max + 1024) / (smem + 1024);
            if (per_sm > 65536 / (warps * 32 * regs)) per_sm = 65536 / (warps * 32 * regs);
            if (per_sm > 2048 / (warps * 32)) per_sm = 2048 / (warps * 32);
            if (per_sm > 32) per_sm = 32;
            if (per_sm < 1) per_sm = 1;
            long long grid = g_env.small_grid_f ? (long long)F : (long long)per_sm * sms;
            if (t.max_clusters > 0 && grid > t.max_clusters) grid = t.max_clusters;
            if (grid > F) grid = F;
            if (grid < 1) grid = 1;
            pl->path = PHNMS_PATH_FUSED;
            pl->variant = PHNMS_FUSED_SMALL;
            pl->cluster = 1;
            pl->threads = warps * 32;
            pl->rows_per_cta = warps * 32;
            pl->smem_bytes = smem;
            pl->grid = (int)grid;
            pl->max_active_clusters = (int)((long long)per_sm * sms);
            return PHNMS_OK;
        }
        if (reg_ok) {
            const int max_cpt = (n_off == 36) ? 2 : 1;
            // The smallest cluster whose CTAs can hold their share of the frame.  Threads: one (n_off 72) or two (n_off 36)
            // proposals each, rounded up to a warp.  A CTA likes >= 8 "spare lanes" beyond its rows (they hold register
            // copies of a fallback batch's candidates, fused_reg.cuh) -- but never at the price of a register-file
            // occupancy step: 128 registers/thread give 4 / 3 / 2 / 1 CTAs per SM at <= 128 / 160 / 256 / 512 threads,
            // and one more warp for the spares at 256 rows halves the resident rows per SM (measured: N = 256: 34 -> 54 M
            // frames/s, N = 512: 18 -> 28 M, N = 2048: 3.4 -> 5.5 M, N = 4096: 1.4 -> 2.3 M without them).
            const int kSpare = 8;
            int csize = 0;
            for (int i = 0; i < 5 && !csize; ++i) {
                const int c = cand[i];
                if (t.cluster && c != t.cluster) continue;
                const int rpc = rows_for(c);
                if (rpc > 512 * max_cpt) continue;
                if (freg_layout(rpc, P, c).total > smem_max) continue;
                csize = c;
            }
            if (csize) {
                const int rpc = rows_for(csize);
                int threads, cpt;
                if (t.threads) {
                    threads = t.threads;
                    if (threads > 512 || threads < 128 || threads % 32) return PHNMS_ERR_TUNING;
                    cpt = (rpc + threads - 1) / threads;
                    if (cpt > max_cpt) return PHNMS_ERR_TUNING;
                } else {
                    cpt = max_cpt;
                    threads = round_up((rpc + cpt - 1) / cpt, 32);
                    if (threads > 512) threads = 512;
                    if (threads < 128) threads = 128;
                    if (threads * cpt < rpc) return PHNMS_ERR_TUNING;
                    if (threads >= rpc) cpt = 1;   // one proposal per thread already covers the rows
                    if (threads * cpt - rpc < kSpare && threads + 32 <= 512 && 512 / (threads + 32) == 512 / threads)
                        threads += 32;             // a warp of spare lanes where it costs no occupancy
                }
                const FregLayout L = freg_layout(rpc, P, csize);
                int per_sm = smem_max / (L.total + 1024);
                const int by_regs = 65536 / (threads * 128);   // __launch_bounds__(512, 1): up to 128 registers per thread
                if (per_sm > by_regs) per_sm = by_regs;
                if (per_sm < 1) per_sm = 1;
                long long resident = (long long)(sms / csize) * per_sm;
                if (resident < 1) resident = 1;
                if (t.max_clusters > 0 && resident > t.max_clusters) resident = t.max_clusters;
                long long clusters = F < resident ? F : resident;
                if (clusters < 1) clusters = 1;
                pl->path = PHNMS_PATH_FUSED;
                pl->variant = PHNMS_FUSED_REG;
                pl->cluster = csize;
                pl->threads = threads;
                pl->cols_per_thread = cpt;
                pl->rows_per_cta = rpc;
                pl->smem_bytes = L.total;
                pl->grid = (int)(clusters * csize);
                pl->launches = small ? 1 : 2;  // (phnms_topm_kernel +) phnms_freg_kernel
                // claim counter + per frame: candidate block (capacity) of the cluster kernel, which also covers the kept-lane
                // block of the streaming path (kStreamMaxK slots), + resume flag and list entry
                pl->workspace_bytes = 1024 + (size_t)F * kTopM * (kHdr + 4 * ((P + 3) & ~3)) + 2 * (((size_t)F * 4 + 255) & ~(size_t)255);
                if (want_stream && k_ok && F < 0x7fffffff) {
                    const StreamShape ss = stream_shape(F, N, n_off, top_k, t, sms, smem_max);
                    if (ss.bad_tuning) return PHNMS_ERR_TUNING;
                    if (ss.ok) {
                        pl->variant = PHNMS_FUSED_STREAM;
                        pl->cluster = 1;
                        pl->threads = ss.warps * 32;
                        pl->cols_per_thread = 1;
                        pl->rows_per_cta = ss.warps * 32;
                        pl->smem_bytes = ss.L.total;
                        pl->grid = ss.grid;
                        // phnms_select_kernel + phnms_stream_kernel + the resume pass (phnms_freg_kernel; not needed when every
                        // proposal of a frame can be drawn)
                        pl->launches = (N <= kSelFullScanN && !t.select_cap && !g_env.select_cap) ? 2 : 3;
                        pl->max_active_clusters = ss.grid;
                    } else if (t.variant == PHNMS_FUSED_STREAM) {
                        return PHNMS_ERR_TUNING;
                    }
                }
                return PHNMS_OK;
            }
            if (t.variant == PHNMS_FUSED_REG || t.variant == PHNMS_FUSED_STREAM) return PHNMS_ERR_TUNING;
        } else if (t.variant == PHNMS_FUSED_REG) {
            return PHNMS_ERR_TUNING;
        }

        // ---- shared-memory-resident variant: any n_off
        int csize = 0, rpc = 0;
        FusedLayout L;
        auto fits = [&](int c) {
            rpc = rows_for(c);
            L = fused_layout(rpc, P, c);
            return L.total <= smem_max;
        };
        if (t.cluster) {
            const int c = t.cluster;
            if (!(c == 1 || c == 2 || c == 4 || c == 8 || c == 16)) return PHNMS_ERR_TUNING;
            if (!fits(c)) return PHNMS_ERR_TUNING;
            csize = c;
        } else {
            for (int i = 0; i < 5 && !csize; ++i)
                if (fits(cand[i])) csize = cand[i];
        }
        if (csize) {
            fits(csize);
            int threads = t.threads ? t.threads : round_up(rpc, 32);
            if (threads > 512) threads = 512;
            if (threads < 128) threads = 128;
            if (threads % 32) return PHNMS_ERR_TUNING;
            int per_sm = smem_max / (L.total + 1024);
            if (per_sm > 2048 / threads) per_sm = 2048 / threads;
            if (per_sm < 1) per_sm = 1;
            if (per_sm > 32) per_sm = 32;
            long long resident = (long long)(sms / csize) * per_sm;
            if (resident < 1) resident = 1;
            if (t.max_clusters > 0 && resident > t.max_clusters) resident = t.max_clusters;
            long long clusters = F < resident ? F : resident;
            if (clusters < 1) clusters = 1;
            pl->path = PHNMS_PATH_FUSED;
            pl->variant = PHNMS_FUSED_SMEM;
            pl->cluster = csize;
            pl->threads = threads;
            pl->rows_per_cta = rpc;
            pl->smem_bytes = L.total;
            pl->grid = (int)(clusters * csize);
            return PHNMS_OK;
        }
        if (t.path == PHNMS_PATH_FUSED) return PHNMS_ERR_TUNING;
    }

    pl->path = PHNMS_PATH_TILED;
    pl->cluster = 1;
    pl->threads = kMaskThreads;
    pl->rows_per_cta = kTile;
    pl->smem_bytes = 2 * kTile * ((((P + 3) & ~3) + 1) * 4 + 8);
    const long long cb = (N + 63) / 64;
    pl->grid = (int)(cb * (cb + 1) / 2);
    pl->launches = 3;
    pl->workspace_bytes = tiled_workspace(F, N);
    return PHNMS_OK;
}

template <typename Kern>
int configure_cluster(Kern kern, const phnms_plan &pl, cudaStream_t stream, cudaLaunchConfig_t *cfg, cudaLaunchAttribute *attr) {
    DeviceInfo d;
    int rc = device_info(&d);
    if (rc) return rc;
    rc = ensure_max_smem(reinterpret_cast<const void *>(kern), d.smem_optin);
    if (rc) return rc;
    *cfg = cudaLaunchConfig_t{};
    cfg->gridDim = dim3((unsigned)pl.grid);
    cfg->blockDim = dim3((unsigned)pl.threads);
    cfg->dynamicSmemBytes = (size_t)pl.smem_bytes;
    cfg->stream = stream;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)pl.cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg->attrs = attr;
    cfg->numAttrs = 1;
    return 0;
}

template <typename Kern, typename... Args>
int launch_cluster(Kern kern, const phnms_plan &pl, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    int rc = configure_cluster(kern, pl, stream, &cfg, attr);
    if (rc) return rc;
    return (int)cudaLaunchKernelEx(&cfg, kern, args...);
}

template <typename Kern>
int occupancy_clusters(Kern kern, phnms_plan pl) {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    pl.grid = pl.cluster * 4096;
    if (configure_cluster(kern, pl, nullptr, &cfg, attr)) return 0;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int fused_occupancy_query(const phnms_plan &pl, int n_off);

// Resident clusters for this launch shape (cudaOccupancyMaxActiveClusters).  The persistent grid must not exceed it:
// a cluster that is not resident only starts when another one has finished ALL of its frames (a second full wave).
// The answer depends only on (device, kernel, cluster, threads, smem); the last one is remembered per host thread.
int fused_occupancy(const phnms_plan &pl, int n_off) {
    if (pl.path != PHNMS_PATH_FUSED) return 0;
    struct Key { int dev, variant, n_off, cluster, threads, smem, cpt, occ; };
    static thread_local Key last = {-1, 0, 0, 0, 0, 0, 0, 0};
    int dev = -1;
    cudaGetDevice(&dev);
    if (last.dev == dev && last.variant == pl.variant && last.n_off == n_off && last.cluster == pl.cluster &&
        last.threads == pl.threads && last.smem == pl.smem_bytes && last.cpt == pl.cols_per_thread)
        return last.occ;
    const int occ = fused_occupancy_query(pl, n_off);
    last = Key{dev, pl.variant, n_off, pl.cluster, pl.threads, pl.smem_bytes, pl.cols_per_thread, occ};
    return occ;
}

void fit_grid_to_occupancy(phnms_plan *pl, int64_t F, int n_off, const phnms_tuning *tun) {
    if (pl->path != PHNMS_PATH_FUSED || pl->variant == PHNMS_FUSED_STREAM || pl->variant == PHNMS_FUSED_SMALL) return;
    const int occ = fused_occupancy(*pl, n_off);
    pl->max_active_clusters = occ;
    if (occ <= 0) return;
    long long clusters = occ;
    if (tun && tun->max_clusters > 0 && clusters > tun->max_clusters) clusters = tun->max_clusters;
    if (clusters > F) clusters = F;
    if (clusters < 1) clusters = 1;
    pl->grid = (int)(clusters * pl->cluster);
}

int fused_occupancy_query(const phnms_plan &pl, int n_off) {
    if (pl.variant == PHNMS_FUSED_REG) {
        if (n_off == 72) return occupancy_clusters(phnms_freg_kernel<72, 1, false, false>, pl);
        if (pl.cols_per_thread == 1) return occupancy_clusters(phnms_freg_kernel<36, 1, false, false>, pl);
        return occupancy_clusters(phnms_freg_kernel<36, 2, false, false>, pl);
    }
    return occupancy_clusters(phnms_fused_kernel, pl);
}

// Compact kept-lane records: one thread per (frame, column) reads keep / num_keep of the call that has just been enqueued
// and stores the value to every destination (local memory or peer GPUs' buffers, then over NVLink).  A separate small
// launch after the NMS kernels (the resume pass may still rewrite a frame's keep / num_keep).  Optionally the SAME launch
// completes the step across GPUs: the last block to finish its stores (device-wide counter) releases this rank's epoch
// flag in every peer (system-scope release, ordered after all record stores of the grid) and then waits for an epoch of
// all ranks -- one launch per step instead of records + flag kernel.
struct CollectArgs {
    int n;
    long long row0;
    long long *dst[kMaxCollectDst];
    // completion (all optional)
    unsigned int *counter;                       // device word, zero between launches (reset by the last block)
    unsigned long long signal_epoch, wait_epoch, timeout_ns;
    unsigned long long *signal_dst[kMaxCollectDst];
    const unsigned long long *wait_src;
    int *status;
    int n_flags;                                 // ranks taking part in the completion
    __device__ int n_sync() const { return n_flags; }
};

__device__ __forceinline__ void peer_signal_and_wait(int t, int n, unsigned long long signal_epoch, unsigned long long *const *signal_dst,
                                                     unsigned long long wait_epoch, const unsigned long long *wait_src,
                                                     unsigned long long timeout_ns, int *status) {
    if (t >= n) return;
    if (signal_epoch) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(signal_dst[t]), "l"(signal_epoch) : "memory");
    }
    if (wait_epoch) {
        unsigned long long t0, now, v;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(wait_src + t) : "memory");
            if (v >= wait_epoch) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (now - t0 > timeout_ns) {
                if (status) atomicExch(status, 1 + t);
                break;
            }
            __nanosleep(200);
        }
    }
}

__global__ void phnms_collect_kernel(const long long *__restrict__ keep, const long long *__restrict__ num, long long F,
                                     int N, int top_k, CollectArgs ca) {
    const int w = top_k + 1;
    // grid-stride: at most one block per SM, so the system-scope fence below (it waits for the block's stores to reach the peers
    // over NVLink) is paid once per SM in parallel, not once per 256 records in two or three waves (measured: ~20 -> ~10 us)
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < F * w; i += (long long)gridDim.x * blockDim.x) {
        const long long f = i / w;
        const int c = (int)(i - f * w);
        const long long cnt = num[f];
        const long long v = c == w - 1 ? cnt : ((c < cnt && c < N) ? keep[f * N + c] : 0ll);
        for (int d = 0; d < ca.n; ++d) ca.dst[d][(ca.row0 + f) * w + c] = v;
    }
    if (ca.counter == nullptr) return;
    // the last block to get here has seen every other block's stores (fence + atomic): it speaks for the grid
    __shared__ int last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(ca.counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    if (threadIdx.x == 0) *ca.counter = 0u;
    peer_signal_and_wait((int)threadIdx.x, ca.n_sync(), ca.signal_epoch, ca.signal_dst, ca.wait_epoch, ca.wait_src, ca.timeout_ns, ca.status);
}

// Cross-GPU completion of a collection step.  Every rank owns a flag array (one u64 per rank) in peer-mapped memory.
// signal: store `epoch` into this rank's slot of every peer's array (system-scope release: ordered after the records the
//         preceding kernels of this stream wrote to the same peers);
// wait:   spin until every slot of the local array has reached `epoch` (system-scope acquire), give up after `timeout_ns`.
struct PeerSyncArgs {
    int n;
    unsigned long long signal_epoch, wait_epoch, timeout_ns;
    unsigned long long *signal_dst[kMaxCollectDst];
    const unsigned long long *wait_src;
    int *status;
};

__global__ void phnms_peer_sync_kernel(PeerSyncArgs a) {
    peer_signal_and_wait((int)threadIdx.x, a.n, a.signal_epoch, a.signal_dst, a.wait_epoch, a.wait_src, a.timeout_ns, a.status);
}

}  // namespace

extern "C" {

int phnms_abi_version(void) { return PHNMS_ABI_VERSION; }

const char *phnms_error_string(int code) {
    switch (code) {
        case PHNMS_OK: return "ok";
        case PHNMS_ERR_BAD_ARG: return "bad argument (null pointer, negative size or misaligned pointer)";
        case PHNMS_ERR_N_OFFSETS: return "Wrong number of offsets: n_off must be in [1, 250]";
        case PHNMS_ERR_TOO_MANY: return "The number of column blocks must be less than MAX_COL_BLOCKS (ceil(N/64) < 1000)";
        case PHNMS_ERR_WORKSPACE: return "workspace missing or too small (see phnms_workspace_bytes)";
        case PHNMS_ERR_DEVICE: return "current CUDA device is not compute capability 10.x (library is sm_100a only)";
        case PHNMS_ERR_TUNING: return "tuning override cannot be honoured for this shape";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "unknown error";
}

size_t phnms_workspace_bytes(int64_t F, int64_t N, int n_off, const phnms_tuning *tuning) {
    phnms_plan pl;
    if (check_shape(F, N, n_off) != PHNMS_OK) return 0;
    if (make_plan(F, N, n_off, -1, tuning, nullptr, &pl) != PHNMS_OK) return 0;
    return pl.workspace_bytes;
}

int phnms_plan_query_topk(int64_t F, int64_t N, int n_off, int64_t top_k, const phnms_tuning *tuning, phnms_plan *plan) {
    if (!plan) return PHNMS_ERR_BAD_ARG;
    int rc = check_shape(F, N, n_off);
    if (rc != PHNMS_OK) return rc;
    DeviceInfo dev;
    const bool have_dev = device_info(&dev) == 0;
    rc = make_plan(F, N, n_off, top_k, tuning, have_dev ? &dev : nullptr, plan);
    if (rc == PHNMS_OK && have_dev && dev.cc_major == 10) fit_grid_to_occupancy(plan, F, n_off, tuning);
    return rc;
}

int phnms_plan_query(int64_t F, int64_t N, int n_off, const phnms_tuning *tuning, phnms_plan *plan) {
    return phnms_plan_query_topk(F, N, n_off, -1, tuning, plan);
}

size_t phnms_order_workspace_bytes(int64_t F, int64_t N) {
    if (F < 0 || N < 0) return 0;
    return (size_t)F * N * 16 + 256;
}

int phnms_order_f32(const float *scores, const int32_t *n_valid, int64_t F, int64_t N, int sort_model, int64_t *order,
                    void *ws, size_t ws_bytes, void *stream) {
    if (F < 0 || N < 0 || sort_model < 0 || sort_model > 2) return PHNMS_ERR_BAD_ARG;
    if (F == 0 || N == 0) return PHNMS_OK;
    if (!scores || !order) return PHNMS_ERR_BAD_ARG;
    if (N > 0x7fffffff / 8) return PHNMS_ERR_TOO_MANY;
    if (!ws || ws_bytes < phnms_order_workspace_bytes(F, N)) return PHNMS_ERR_WORKSPACE;
    uint32_t *w = reinterpret_cast<uint32_t *>(((uintptr_t)ws + 255) & ~(uintptr_t)255);
    phnms_order_kernel<<<(unsigned)F, kOrderThreads, 0, (cudaStream_t)stream>>>(
        scores, n_valid, (int)N, sort_model, reinterpret_cast<long long *>(order), w);
    return (int)cudaGetLastError();
}

static int forward_impl(const float *props, const float *scores, const int32_t *n_valid, int64_t F, int64_t N, int n_off,
                        float thresh, int64_t top_k, int sort_model, int64_t *keep, int64_t *num_keep, int64_t *parent,
                        void *ws, size_t ws_bytes, const phnms_tuning *tuning, void *stream_, int64_t *trace, int trace_len,
                        const RecordSink *rec = nullptr, bool *rec_done = nullptr);

int phnms_forward_f32(const float *props, const float *scores, const int32_t *n_valid, int64_t F, int64_t N, int n_off,
                      float thresh, int64_t top_k, int sort_model, int64_t *keep, int64_t *num_keep, int64_t *parent,
                      void *ws, size_t ws_bytes, const phnms_tuning *tuning, void *stream_) {
    return forward_impl(props, scores, n_valid, F, N, n_off, thresh, top_k, sort_model, keep, num_keep, parent, ws, ws_bytes,
                        tuning, stream_, nullptr, 0);
}

int phnms_forward_f32_trace(const float *props, const float *scores, const int32_t *n_valid, int64_t F, int64_t N,
                            int n_off, float thresh, int64_t top_k, int sort_model, int64_t *keep, int64_t *num_keep,
                            int64_t *parent, void *ws, size_t ws_bytes, const phnms_tuning *tuning, void *stream_,
                            int64_t *trace, int trace_len) {
    return forward_impl(props, scores, n_valid, F, N, n_off, thresh, top_k, sort_model, keep, num_keep, parent, ws, ws_bytes,
                        tuning, stream_, trace, trace_len);
}

int phnms_forward_collect_f32(const float *props, const float *scores, const int32_t *n_valid, int64_t F, int64_t N,
                              int n_off, float thresh, int64_t top_k, int sort_model, int64_t *keep, int64_t *num_keep,
                              int64_t *parent, void *ws, size_t ws_bytes, const phnms_tuning *tuning, void *stream_,
                              const phnms_collect *collect) {
    if (!collect) return PHNMS_ERR_BAD_ARG;
    if (collect->n_dst < 1 || collect->n_dst > PHNMS_MAX_DST || collect->row0 < 0 || top_k < 1 || top_k > 0x7ffffff)
        return PHNMS_ERR_BAD_ARG;
    // a record is never stored outside a destination: the buffers are [rows, width] and this call fills rows row0 .. row0+F
    if ((int64_t)collect->width != top_k + 1 || F < 0 || collect->row0 + F > collect->rows) return PHNMS_ERR_BAD_ARG;
    for (int d = 0; d < collect->n_dst; ++d)
        if (!collect->dst[d] || ((uintptr_t)collect->dst[d] & 7u)) return PHNMS_ERR_BAD_ARG;
    RecordSink sink = {};
    sink.n = collect->n_dst;
    sink.width = (int)(top_k + 1);
    sink.row0 = collect->row0;
    for (int d = 0; d < collect->n_dst; ++d) sink.dst[d] = reinterpret_cast<long long *>(collect->dst[d]);
    bool rec_done = false;
    int rc = forward_impl(props, scores, n_valid, F, N, n_off, thresh, top_k, sort_model, keep, num_keep, parent, ws,
                          ws_bytes, tuning, stream_, nullptr, 0, N > 0 ? &sink : nullptr, &rec_done);
    if (rc != PHNMS_OK || F == 0) return rc;
    if (rec_done) {
        // The NMS kernels stored the records themselves.  What is left is the completion across GPUs: one single-block launch
        // (system-scope release of this rank's epoch after the kernels above, acquire of the epoch the consumer needs).
        if (!collect->signal_epoch && !collect->wait_epoch) return PHNMS_OK;
        if (collect->wait_epoch && !collect->wait_src) return PHNMS_ERR_BAD_ARG;
        PeerSyncArgs a;
        a.n = collect->n_dst;
        a.signal_epoch = collect->signal_epoch;
        a.wait_epoch = collect->wait_epoch;
        a.timeout_ns = collect->timeout_ns ? collect->timeout_ns : 10000000000ull;
        a.wait_src = reinterpret_cast<const unsigned long long *>(collect->wait_src);
        a.status = collect->status;
        for (int d = 0; d < kMaxCollectDst; ++d) a.signal_dst[d] = nullptr;
        for (int d = 0; d < collect->n_dst; ++d) {
            if (collect->signal_epoch && (!collect->signal_dst[d] || ((uintptr_t)collect->signal_dst[d] & 7u))) return PHNMS_ERR_BAD_ARG;
            a.signal_dst[d] = reinterpret_cast<unsigned long long *>(collect->signal_dst[d]);
        }
        phnms_peer_sync_kernel<<<1, 32, 0, (cudaStream_t)stream_>>>(a);
        return (int)cudaGetLastError();
    }
    CollectArgs ca;
    ca.n = collect->n_dst;
    ca.row0 = collect->row0;
    for (int d = 0; d < kMaxCollectDst; ++d) ca.dst[d] = d < collect->n_dst ? reinterpret_cast<long long *>(collect->dst[d]) : nullptr;
    ca.counter = nullptr;
    ca.signal_epoch = ca.wait_epoch = 0ull;
    ca.timeout_ns = 0ull;
    ca.wait_src = nullptr;
    ca.status = nullptr;
    ca.n_flags = 0;
    for (int d = 0; d < kMaxCollectDst; ++d) ca.signal_dst[d] = nullptr;
    if (collect->signal_epoch || collect->wait_epoch) {   // completion across GPUs folded into the same launch
        if (!collect->sync_counter || ((uintptr_t)collect->sync_counter & 3u)) return PHNMS_ERR_BAD_ARG;
        if (collect->wait_epoch && !collect->wait_src) return PHNMS_ERR_BAD_ARG;
        ca.counter = collect->sync_counter;
        ca.signal_epoch = collect->signal_epoch;
        ca.wait_epoch = collect->wait_epoch;
        ca.timeout_ns = collect->timeout_ns ? collect->timeout_ns : 10000000000ull;
        ca.wait_src = reinterpret_cast<const unsigned long long *>(collect->wait_src);
        ca.status = collect->status;
        ca.n_flags = collect->n_dst;
        for (int d = 0; d < collect->n_dst; ++d) {
            if (collect->signal_epoch && (!collect->signal_dst[d] || ((uintptr_t)collect->signal_dst[d] & 7u))) return PHNMS_ERR_BAD_ARG;
            ca.signal_dst[d] = reinterpret_cast<unsigned long long *>(collect->signal_dst[d]);
        }
    }
    const long long total = (long long)F * (top_k + 1);
    DeviceInfo dinfo;
    if (device_info(&dinfo) != 0) dinfo.sms = 148;
    const long long blocks = (total + 255) / 256;
    phnms_collect_kernel<<<(unsigned)(blocks < dinfo.sms ? blocks : dinfo.sms), 256, 0, (cudaStream_t)stream_>>>(
        reinterpret_cast<const long long *>(keep), reinterpret_cast<const long long *>(num_keep), F, (int)N, (int)top_k, ca);
    return (int)cudaGetLastError();
}

// ---- the streaming path: select -> stream -> resume ----------------------------------------------------------------------
static int launch_stream(const float *props, const float *scores, const int32_t *n_valid, int64_t F, int64_t N, int n_off,
                         float thresh, int64_t top_k, int sort_model, int64_t *keep, int64_t *num_keep, int64_t *parent,
                         void *ws, size_t ws_bytes, const phnms_tuning *tuning, const DeviceInfo &dev, const phnms_plan &pl,
                         cudaStream_t stream, const RecordSink &rec) {
    if (!ws || ws_bytes < pl.workspace_bytes) return PHNMS_ERR_WORKSPACE;
    const phnms_tuning t = tuning ? *tuning : phnms_tuning{};
    const int P = 5 + n_off, SLOT = kHdr + 4 * ((P + 3) & ~3);
    const StreamShape ss = stream_shape(F, N, n_off, top_k, t, dev.sms, dev.smem_optin);
    if (!ss.ok) return fail_at("stream shape", PHNMS_ERR_TUNING);
    unsigned char *base = reinterpret_cast<unsigned char *>(((uintptr_t)ws + 255) & ~(uintptr_t)255);
    unsigned int *ctrs = reinterpret_cast<unsigned int *>(base + 64);          // (the cluster kernel's claim counter sits at +0)
    unsigned char *blocks = base + 256;
    unsigned char *after = blocks + (((size_t)F * kTopM * SLOT + 255) & ~(size_t)255);
    int *flags = reinterpret_cast<int *>(after);
    int *list = reinterpret_cast<int *>(after + (((size_t)F * 4 + 255) & ~(size_t)255));

    // Draws per frame before the select kernel hands a frame over (open).  Frames of up to 256 proposals -- everything PHNet
    // itself produces (240 priors) -- are always scanned to the end: no frame is ever left open and the resume pass is not
    // even launched.  Larger frames: 64, and 128 when more than four lanes are wanted (deeper draws cost every frame that has fewer
    // lanes than top_k; measured at top_k = 8: 2.3 % of the generator's frames need a lane beyond draw 64 and pay the resume pass).
    int cap = t.select_cap ? t.select_cap : (g_env.select_cap ? g_env.select_cap : (N <= kSelFullScanN ? (int)N : (top_k > 4 ? 2 * kSelCapDefault : kSelCapDefault)));
    if (cap < kSelBatch) {
        if (t.select_cap || g_env.select_cap) return PHNMS_ERR_TUNING;   // an explicit cap below one batch
        cap = kSelBatch;   // frames of fewer than 8 proposals: one batch draws them all
    }
    if (!(g_env.skip & 1)) {   // (1) the greedy scan over the best-ranked proposals: one warp per frame
        SelectParams sp;
        sp.props = props; sp.scores = scores; sp.n_valid = n_valid;
        sp.F = F; sp.N = (int)N; sp.n_off = n_off; sp.sort_model = sort_model; sp.top_k = (int)top_k; sp.cap = cap;
        sp.thr = thresh;
        sp.blocks = blocks; sp.block_bytes = ss.block_bytes;
        sp.flags = flags; sp.ctrs = ctrs;
        sp.keep = reinterpret_cast<long long *>(keep);
        sp.num_keep = reinterpret_cast<long long *>(num_keep);
        int warps = kSelWarps;
        while (warps > 1 && select_smem_bytes((int)N, n_off, (int)top_k, warps) > 100 * 1024) warps >>= 1;
        const size_t sm = select_smem_bytes((int)N, n_off, (int)top_k, warps);
        // (static + dynamic shared memory beyond 48 KB needs the opt-in; cached per device, so simply always)
        if (rec.n > 0) {
            const int e2 = ensure_max_smem(reinterpret_cast<const void *>(phnms_select_kernel<true>), dev.smem_optin);
            if (e2) return fail_at("select smem attribute", e2);
            phnms_select_kernel<true><<<(unsigned)((F + warps - 1) / warps), warps * 32, sm, stream>>>(sp, rec);
        } else {
            const int e2 = ensure_max_smem(reinterpret_cast<const void *>(phnms_select_kernel<false>), dev.smem_optin);
            if (e2) return fail_at("select smem attribute", e2);
            phnms_select_kernel<false><<<(unsigned)((F + warps - 1) / warps), warps * 32, sm, stream>>>(sp, rec);
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail_at("select launch", (int)e);
    }
    if (!(g_env.skip & 2)) {   // (2) every proposal against its frame's kept lanes
        StreamParams q;
        q.props = props; q.scores = scores; q.n_valid = n_valid;
        q.keep = reinterpret_cast<long long *>(keep);
        q.parent = reinterpret_cast<long long *>(parent);
        q.blocks = blocks; q.block_bytes = ss.block_bytes;
        q.flags = flags; q.ctrs = ctrs; q.list = list;
        q.F = F; q.N = (int)N; q.top_k = (int)top_k; q.sort_model = sort_model; q.thr = thresh;
        q.ipf = ss.ipf; q.nseg = ss.nseg; q.ips = ss.ips; q.ks = ss.ks; q.bundle = ss.bundle;
        q.off_ring = ss.L.off_ring; q.off_slots = ss.L.off_slots; q.slot_bytes = ss.L.slot_bytes; q.off_bit = ss.L.off_bit;
        const int lanes = ss.lanes;
        int rc = 0;
// (Measured and rejected: launching the streaming kernel with programmatic stream serialization -- griddepcontrol.launch_dependents at
// the top of the select kernel, griddepcontrol.wait before the first kept-block request -- so that its prologue overlaps the select
// kernel's tail: 0.835 -> 0.724 of the roofline at the headline shape, no effect at 36 offsets.)
#define PHNMS_LAUNCH_STREAM(NO, NK, GE)                                                                               \
    do {                                                                                                              \
        rc = ensure_max_smem(reinterpret_cast<const void *>(phnms_stream_kernel<NO, NK, GE>), dev.smem_optin);        \
        if (rc) return fail_at("stream smem attribute", rc);                                                          \
        phnms_stream_kernel<NO, NK, GE><<<(unsigned)ss.grid, ss.warps * 32, (size_t)ss.L.total, stream>>>(q);         \
    } while (0)
#define PHNMS_LAUNCH_STREAM_NO(NO)                                                                                    \
    do {                                                                                                              \
        if (general) {                                                                                                \
            if (lanes == 4) PHNMS_LAUNCH_STREAM(NO, 4, true);                                                         \
            else if (lanes == 2) PHNMS_LAUNCH_STREAM(NO, 2, true);                                                    \
            else PHNMS_LAUNCH_STREAM(NO, 1, true);                                                                    \
        } else {                                                                                                      \
            if (lanes == 4) PHNMS_LAUNCH_STREAM(NO, 4, false);                                                        \
            else if (lanes == 2) PHNMS_LAUNCH_STREAM(NO, 2, false);                                                   \
            else PHNMS_LAUNCH_STREAM(NO, 1, false);                                                                   \
        }                                                                                                             \
    } while (0)
        const bool general = ss.nseg != 1 || ss.bundle != 1 || n_valid != nullptr;
        if (n_off == 72) PHNMS_LAUNCH_STREAM_NO(72);
        else PHNMS_LAUNCH_STREAM_NO(36);
#undef PHNMS_LAUNCH_STREAM_NO
#undef PHNMS_LAUNCH_STREAM
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail_at("stream launch", (int)e);
    }
    if (N <= cap || (g_env.skip & 4)) return PHNMS_OK;   // every proposal can be drawn: the select kernel never leaves a frame open
    // (3) resume: frames the select kernel left open AND the streaming pass found unfinished are redone from scratch by the
    // register-resident cluster kernel (in-kernel candidate selection), working through the device-side list
    phnms_tuning rt = phnms_tuning{};
    rt.path = PHNMS_PATH_FUSED;
    rt.variant = PHNMS_FUSED_REG;
    rt.max_clusters = t.max_clusters;
    phnms_plan rp;
    int rc = make_plan(F, N, n_off, top_k, &rt, &dev, &rp);
    if (rc != PHNMS_OK) return fail_at("resume plan", rc);
    if (F > 1) fit_grid_to_occupancy(&rp, F, n_off, &rt);
    FusedParams fp;
    fp.props = props; fp.scores = scores; fp.n_valid = n_valid;
    fp.keep = reinterpret_cast<long long *>(keep);
    fp.num_keep = reinterpret_cast<long long *>(num_keep);
    fp.parent = reinterpret_cast<long long *>(parent);
    fp.F = F; fp.top_k = top_k; fp.N = (int)N; fp.n_off = n_off;
    fp.rpc = rp.rows_per_cta; fp.csize = rp.cluster; fp.sort_model = sort_model; fp.thr = thresh;
    fp.L = fused_layout(rp.rows_per_cta, P, rp.cluster);
    fp.trace = nullptr; fp.trace_len = 0;
    fp.topm = nullptr; fp.topm_count = 0; fp.claim_ctr = nullptr;
    fp.frame_list = list; fp.frame_count = ctrs;
    fp.rec = rec;   // (a frame that is redone gets its record rewritten)
    const FregLayout RL = freg_layout(rp.rows_per_cta, P, rp.cluster);
    if (n_off == 72) return fail_at("resume launch", launch_cluster(phnms_freg_kernel<72, 1, false, false>, rp, stream, fp, RL));
    if (rp.cols_per_thread == 1) return fail_at("resume launch", launch_cluster(phnms_freg_kernel<36, 1, false, false>, rp, stream, fp, RL));
    return fail_at("resume launch", launch_cluster(phnms_freg_kernel<36, 2, false, false>, rp, stream, fp, RL));
}

static int forward_impl(const float *props, const float *scores, const int32_t *n_valid, int64_t F, int64_t N, int n_off,
                        float thresh, int64_t top_k, int sort_model, int64_t *keep, int64_t *num_keep, int64_t *parent,
                        void *ws, size_t ws_bytes, const phnms_tuning *tuning, void *stream_, int64_t *trace, int trace_len,
                        const RecordSink *rec_in, bool *rec_done) {
    // compact records: written by the kernels themselves on the paths that can (streaming, small-frame, register-resident
    // cluster kernel); *rec_done tells the caller whether it still has to launch the separate record kernel
    RecordSink rec = {};
    if (rec_in && !g_env.no_fused_records && !trace) rec = *rec_in;
    if (rec_done) *rec_done = false;
    int rc = check_shape(F, N, n_off);
    if (rc != PHNMS_OK) return rc;
    if (sort_model < 0 || sort_model > 2 || top_k < 0) return PHNMS_ERR_BAD_ARG;
    if (F == 0) return PHNMS_OK;
    if (!num_keep) return PHNMS_ERR_BAD_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (N == 0) return (int)cudaMemsetAsync(num_keep, 0, (size_t)F * 8, stream);
    if (!props || !scores || !keep || !parent) return PHNMS_ERR_BAD_ARG;
    if (((uintptr_t)props | (uintptr_t)scores) & 3u) return PHNMS_ERR_BAD_ARG;
    if (((uintptr_t)keep | (uintptr_t)parent | (uintptr_t)num_keep) & 7u) return PHNMS_ERR_BAD_ARG;

    DeviceInfo dev;
    rc = device_info(&dev);
    if (rc != 0) return rc;
    if (dev.cc_major != 10) return PHNMS_ERR_DEVICE;
    phnms_plan pl;
    phnms_tuning traced;
    if (trace) {   // the phase trace belongs to the register-resident cluster kernel
        traced = tuning ? *tuning : phnms_tuning{};
        if (traced.variant == 0 || traced.variant == PHNMS_FUSED_STREAM || traced.variant == PHNMS_FUSED_SMALL) traced.variant = PHNMS_FUSED_REG;
        tuning = &traced;
    }
    rc = make_plan(F, N, n_off, top_k, tuning, &dev, &pl);
    if (rc != PHNMS_OK) return rc;
    if (F > 1) fit_grid_to_occupancy(&pl, F, n_off, tuning);
    if (pl.path == PHNMS_PATH_FUSED && pl.variant == PHNMS_FUSED_SMALL && !trace) {   // one launch, one CTA per frame
        SmallParams q;
        q.props = props; q.scores = scores; q.n_valid = n_valid;
        q.keep = reinterpret_cast<long long *>(keep);
        q.num_keep = reinterpret_cast<long long *>(num_keep);
        q.parent = reinterpret_cast<long long *>(parent);
        q.F = F; q.top_k = top_k; q.N = (int)N; q.sort_model = sort_model; q.thr = thresh;
        if (rec_done) *rec_done = rec.n > 0;
#define PHNMS_LAUNCH_SMALL_R(NO, MT, MB, RE)                                                                            \
    do {                                                                                                              \
        rc = ensure_max_smem(reinterpret_cast<const void *>(phnms_small_kernel<NO, MT, MB, RE>), dev.smem_optin);     \
        if (rc) return fail_at("small smem attribute", rc);                                                           \
        phnms_small_kernel<NO, MT, MB, RE><<<(unsigned)pl.grid, pl.threads, (size_t)pl.smem_bytes, stream>>>(q, rec); \
    } while (0)
#define PHNMS_LAUNCH_SMALL(NO, MT, MB)                                                                                  \
    do {                                                                                                              \
        if (rec.n > 0) PHNMS_LAUNCH_SMALL_R(NO, MT, MB, true);                                                        \
        else PHNMS_LAUNCH_SMALL_R(NO, MT, MB, false);                                                                 \
    } while (0)
        if (n_off == 72) PHNMS_LAUNCH_SMALL(72, 512, 1);
        else if (pl.threads <= 256) PHNMS_LAUNCH_SMALL(36, 256, 3);
        else PHNMS_LAUNCH_SMALL(36, 512, 1);
#undef PHNMS_LAUNCH_SMALL
#undef PHNMS_LAUNCH_SMALL_R
        return fail_at("small launch", (int)cudaGetLastError());
    }
    if (pl.path == PHNMS_PATH_FUSED && pl.variant == PHNMS_FUSED_STREAM) {
        if (rec_done) *rec_done = rec.n > 0 && !g_env.skip;
        return launch_stream(props, scores, n_valid, F, N, n_off, thresh, top_k, sort_model, keep, num_keep, parent, ws, ws_bytes,
                             tuning, dev, pl, stream, rec);
    }

    if (pl.path == PHNMS_PATH_FUSED) {
        FusedParams fp;
        fp.props = props;
        fp.scores = scores;
        fp.n_valid = n_valid;
        fp.keep = reinterpret_cast<long long *>(keep);
        fp.num_keep = reinterpret_cast<long long *>(num_keep);
        fp.parent = reinterpret_cast<long long *>(parent);
        fp.F = F;
        fp.top_k = top_k;
        fp.N = (int)N;
        fp.n_off = n_off;
        fp.rpc = pl.rows_per_cta;
        fp.csize = pl.cluster;
        fp.sort_model = sort_model;
        fp.thr = thresh;
        fp.L = fused_layout(pl.rows_per_cta, 5 + n_off, pl.cluster);
        fp.trace = reinterpret_cast<long long *>(trace);
        fp.trace_len = trace_len;
        fp.topm = nullptr;
        fp.topm_count = 0;
        fp.claim_ctr = nullptr;
        fp.frame_list = nullptr;
        fp.frame_count = nullptr;
        fp.rec = RecordSink{};
        if (pl.variant == PHNMS_FUSED_REG) {
            if (!ws || ws_bytes < pl.workspace_bytes) return PHNMS_ERR_WORKSPACE;
            fp.rec = rec;
            if (rec_done) *rec_done = rec.n > 0;
            unsigned long long *claim_ctr = reinterpret_cast<unsigned long long *>(((uintptr_t)ws + 255) & ~(uintptr_t)255);
            int *topm = reinterpret_cast<int *>(reinterpret_cast<unsigned char *>(claim_ctr) + 256);
            {   // frames are claimed dynamically by default only where that measured faster (single-CTA frames, one
                // proposal per thread: +2.5 %); clusters pay ~3 % for the split barrier, the 2-proposal kernel ~7 %
                const int sched = tuning ? tuning->schedule : 0;
                const bool dynamic = sched == PHNMS_SCHED_DYNAMIC || (sched == 0 && pl.cluster == 1 && pl.cols_per_thread == 1);
                fp.claim_ctr = dynamic ? claim_ctr : nullptr;
            }
            // enough candidates that the first batch usually reaches top_k without an exchange
            // Candidates per frame.  With top_k <= 4 eight of them leave ~14 % of the frames short of top_k kept lanes (those
            // frames pay a fallback batch: cluster exchange, barriers), twelve ~3 %, sixteen ~0.5 % -- but every candidate
            // costs time in the top-M kernel (one warp per frame) whatever the frame size.  Measured: N = 1000 x 72:
            // 14.9 / 15.6 / 15.4 M frames/s with 8 / 12 / 16; N = 240: 53.9 / 51.4 / 48.1 M.
            int topm_count = (top_k > 0 && top_k <= 4) ? (N > 384 ? 12 : 8) : kTopM;
            if (g_env.topm_count >= 2 && g_env.topm_count <= kTopM) topm_count = g_env.topm_count;   // experiment knob
            fp.topm_count = topm_count;
            const bool single_launch = pl.launches == 1 || g_env.no_topm;   // small calls: in-kernel selection, static schedule
            if (single_launch) {
                fp.claim_ctr = nullptr;   // (the claim counter is zeroed by the top-M kernel, which does not run)
                fp.topm_count = 0;
            } else {
                int warps = kTopmWarps;
                while (warps > 1 && topm_smem_bytes((int)N, warps) > 160 * 1024) warps >>= 1;
                const size_t sm = topm_smem_bytes((int)N, warps);
                {
                    const int e2 = ensure_max_smem(reinterpret_cast<const void *>(phnms_topm_kernel), dev.smem_optin);
                    if (e2) return e2;
                }
                phnms_topm_kernel<<<(unsigned)((F + warps - 1) / warps), warps * 32, sm, stream>>>(
                    props, scores, n_valid, F, (int)N, n_off, sort_model, topm_count, thresh, topm, claim_ctr, top_k,
                    reinterpret_cast<long long *>(keep));
            }
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) return (int)e;
            fp.topm = single_launch ? nullptr : topm;
            const FregLayout RL = freg_layout(pl.rows_per_cta, 5 + n_off, pl.cluster);
#define PHNMS_LAUNCH_FREG(TR, DY)                                                                              \
    do {                                                                                                       \
        if (n_off == 72) return launch_cluster(phnms_freg_kernel<72, 1, TR, DY>, pl, stream, fp, RL);            \
        if (pl.cols_per_thread == 1) return launch_cluster(phnms_freg_kernel<36, 1, TR, DY>, pl, stream, fp, RL); \
        return launch_cluster(phnms_freg_kernel<36, 2, TR, DY>, pl, stream, fp, RL);                            \
    } while (0)
            const bool dynamic = fp.claim_ctr != nullptr;
            if (trace) {   // profiling / watchdog build of the same kernel
                if (dynamic) PHNMS_LAUNCH_FREG(true, true);
                PHNMS_LAUNCH_FREG(true, false);
            }
            if (dynamic) PHNMS_LAUNCH_FREG(false, true);
            PHNMS_LAUNCH_FREG(false, false);
#undef PHNMS_LAUNCH_FREG
        }
        return launch_cluster(phnms_fused_kernel, pl, stream, fp);
    }

    // ---- tiled path: order -> bitmask -> scan ------------------------------------------------------------
    if (!ws || ws_bytes < pl.workspace_bytes) return PHNMS_ERR_WORKSPACE;
    unsigned char *base = reinterpret_cast<unsigned char *>(((uintptr_t)ws + 255) & ~(uintptr_t)255);
    long long *order = reinterpret_cast<long long *>(base);
    uint32_t *sort_ws = reinterpret_cast<uint32_t *>(base + (size_t)F * N * 8);
    unsigned long long *mask = reinterpret_cast<unsigned long long *>(base + (size_t)F * N * 24);
    const int col_blocks = (int)((N + 63) / 64);
    const int P = 5 + n_off;

    phnms_order_kernel<<<(unsigned)F, kOrderThreads, 0, stream>>>(scores, n_valid, (int)N, sort_model, order, sort_ws);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;

    e = cudaFuncSetAttribute(phnms_mask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem_bytes);
    if (e != cudaSuccess) return (int)e;
    // gridDim.y is limited to 65535 frames per launch
    for (int64_t f0 = 0; f0 < F; f0 += 65535) {
        const int64_t fc = (F - f0) < 65535 ? (F - f0) : 65535;
        dim3 grid((unsigned)pl.grid, (unsigned)fc);
        phnms_mask_kernel<<<grid, kMaskThreads, (size_t)pl.smem_bytes, stream>>>(
            props + (size_t)f0 * N * P, order + (size_t)f0 * N, n_valid ? n_valid + f0 : nullptr, (int)N, n_off, thresh,
            col_blocks, mask + (size_t)f0 * N * col_blocks, props + (size_t)F * N * P);
        e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
    }
    phnms_scan_kernel<<<(unsigned)F, 32, 0, stream>>>(order, mask, n_valid, (int)N, col_blocks, top_k,
                                                      reinterpret_cast<long long *>(keep),
                                                      reinterpret_cast<long long *>(num_keep),
                                                      reinterpret_cast<long long *>(parent));
    return (int)cudaGetLastError();
}

// ---- get_lanes for a clip: prepare -> lane NMS -> gather -----------------------------------------------------------------
static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// get_lanes as ONE launch (frontend.cuh): the offset counts PHNet ships, at most 1024 priors, top_k within the select step's range
static bool get_lanes_fused_ok(int64_t A, int n_off, int64_t top_k, const phnms_tuning *tuning) {
    if (tuning && (tuning->path || tuning->variant || tuning->cluster || tuning->threads)) return false;   // an explicit device path
    return (n_off == 36 || n_off == 72) && A <= 1024 && (top_k < 0 || (top_k >= 1 && top_k <= kStreamMaxK));
}

size_t phnms_get_lanes_workspace_bytes(int64_t T, int64_t A, int n_off, const phnms_tuning *tuning) {
    if (T < 0 || A < 0 || check_shape(T, A, n_off) != PHNMS_OK) return 0;
    const size_t P = 5 + (size_t)n_off;
    return 256 + align256((size_t)T * A * P * 4) + align256((size_t)T * A * 4) * 2 + align256((size_t)T * 4) +
           align256((size_t)T * A * 8) * 2 + align256((size_t)T * 8) + align256(phnms_workspace_bytes(T, A, n_off, tuning));
}

int phnms_get_lanes_f32(const float *pred, int64_t T, int64_t A, int n_off, int hdr, float conf_threshold, float img_w,
                        float nms_thres, int64_t top_k, int sort_model, float *out_rows, int64_t *out_num,
                        int64_t *out_index, unsigned char *keep_inds, void *ws, size_t ws_bytes,
                        const phnms_tuning *tuning, void *stream_) {
    int rc = check_shape(T, A, n_off);
    if (rc != PHNMS_OK) return rc;
    if ((hdr != 6 && hdr != 7) || top_k < 1 || top_k > A) return PHNMS_ERR_BAD_ARG;
    if (T == 0 || A == 0) return PHNMS_OK;
    if (!pred || !out_rows || !out_num || !out_index || !keep_inds) return PHNMS_ERR_BAD_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (get_lanes_fused_ok(A, n_off, top_k, tuning) && !g_env.no_stream) {   // one launch, no workspace
        DeviceInfo dev;
        rc = device_info(&dev);
        if (rc != 0) return rc;
        if (dev.cc_major != 10) return PHNMS_ERR_DEVICE;
        GetLanesFusedParams gp;
        gp.pred = pred; gp.T = T; gp.A = (int)A; gp.hdr = hdr; gp.n_off = n_off; gp.top_k = (int)top_k; gp.sort_model = sort_model;
        gp.conf_thr = conf_threshold; gp.img_w_m1 = img_w - 1.0f; gp.n_strips = (float)(n_off - 1); gp.thr = nms_thres;
        gp.out_rows = out_rows; gp.out_num = reinterpret_cast<long long *>(out_num);
        gp.out_index = reinterpret_cast<long long *>(out_index); gp.keep_inds = keep_inds;
        const int warps = kSelWarps;
        const size_t sm = (size_t)warps * get_lanes_warp_words((int)A, n_off, (int)top_k) * 4;
        rc = ensure_max_smem(reinterpret_cast<const void *>(phnms_get_lanes_fused_kernel), dev.smem_optin);
        if (rc) return rc;
        phnms_get_lanes_fused_kernel<<<(unsigned)((T + warps - 1) / warps), warps * 32, sm, stream>>>(gp);
        return (int)cudaGetLastError();
    }
    if (!ws || ws_bytes < phnms_get_lanes_workspace_bytes(T, A, n_off, tuning)) return PHNMS_ERR_WORKSPACE;
    const size_t P = 5 + (size_t)n_off;
    unsigned char *b = reinterpret_cast<unsigned char *>(align256((size_t)ws));
    float *cprops = reinterpret_cast<float *>(b);            b += align256((size_t)T * A * P * 4);
    float *cscores = reinterpret_cast<float *>(b);           b += align256((size_t)T * A * 4);
    int *src = reinterpret_cast<int *>(b);                   b += align256((size_t)T * A * 4);
    int *n_valid = reinterpret_cast<int *>(b);               b += align256((size_t)T * 4);
    int64_t *keep = reinterpret_cast<int64_t *>(b);          b += align256((size_t)T * A * 8);
    int64_t *parent = reinterpret_cast<int64_t *>(b);        b += align256((size_t)T * A * 8);
    unsigned char *nms_ws = b;
    const size_t nms_ws_bytes = phnms_workspace_bytes(T, A, n_off, tuning);
    const float n_strips = (float)(n_off - 1);
    phnms_prepare_kernel<<<(unsigned)T, kPrepThreads, (size_t)A * sizeof(int), stream>>>(
        pred, (int)A, hdr, n_off, conf_threshold, img_w - 1.0f, n_strips, cprops, cscores, src, n_valid, keep_inds);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    rc = phnms_forward_f32(cprops, cscores, n_valid, T, A, n_off, nms_thres, top_k, sort_model, keep, out_num, parent,
                           nms_ws_bytes ? nms_ws : nullptr, nms_ws_bytes, tuning, stream_);
    if (rc != PHNMS_OK) return rc;
    phnms_gather_kernel<<<(unsigned)T, 128, 0, stream>>>(pred, (int)A, hdr, n_off, n_strips,
                                                       reinterpret_cast<const long long *>(keep),
                                                       reinterpret_cast<const long long *>(out_num), src, (int)top_k,
                                                       out_rows, reinterpret_cast<long long *>(out_index));
    return (int)cudaGetLastError();
}

size_t phnms_ordered_f64_workspace_bytes(int64_t F, int64_t N) {
    if (F < 0 || N < 0) return 0;
    return (size_t)F * N * (size_t)((N + 63) / 64) * 8 + 256;
}

int phnms_forward_ordered_f64(const double *props, const int64_t *order, const int32_t *n_valid, int64_t F, int64_t N,
                              int n_off, float thresh, int64_t top_k, int64_t *keep, int64_t *num_keep, int64_t *parent,
                              void *ws, size_t ws_bytes, void *stream_) {
    int rc = check_shape(F, N, n_off);
    if (rc != PHNMS_OK) return rc;
    if (top_k < 0) return PHNMS_ERR_BAD_ARG;
    if (F == 0) return PHNMS_OK;
    if (!num_keep) return PHNMS_ERR_BAD_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (N == 0) return (int)cudaMemsetAsync(num_keep, 0, (size_t)F * 8, stream);
    if (!props || !order || !keep || !parent) return PHNMS_ERR_BAD_ARG;
    if (((uintptr_t)props | (uintptr_t)order | (uintptr_t)keep | (uintptr_t)parent | (uintptr_t)num_keep) & 7u)
        return PHNMS_ERR_BAD_ARG;
    if (!ws || ws_bytes < phnms_ordered_f64_workspace_bytes(F, N)) return PHNMS_ERR_WORKSPACE;
    DeviceInfo dev;
    rc = device_info(&dev);
    if (rc != 0) return rc;
    if (dev.cc_major != 10) return PHNMS_ERR_DEVICE;
    unsigned long long *mask = reinterpret_cast<unsigned long long *>(((uintptr_t)ws + 255) & ~(uintptr_t)255);
    const int col_blocks = (int)((N + 63) / 64);
    const long long *ord = reinterpret_cast<const long long *>(order);
    for (int64_t f0 = 0; f0 < F; f0 += 65535) {            // gridDim.z is limited to 65535 frames per launch
        const int64_t fc = (F - f0) < 65535 ? (F - f0) : 65535;
        dim3 grid((unsigned)((N + 127) / 128), (unsigned)col_blocks, (unsigned)fc);
        phnms_mask_f64_kernel<<<grid, 128, 0, stream>>>(props + (size_t)f0 * N * (5 + n_off), ord + (size_t)f0 * N,
                                                        n_valid ? n_valid + f0 : nullptr, (int)N, n_off, thresh,
                                                        col_blocks, mask + (size_t)f0 * N * col_blocks);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
    }
    phnms_scan_kernel<<<(unsigned)F, 32, 0, stream>>>(ord, mask, n_valid, (int)N, col_blocks, top_k,
                                                      reinterpret_cast<long long *>(keep),
                                                      reinterpret_cast<long long *>(num_keep),
                                                      reinterpret_cast<long long *>(parent));
    return (int)cudaGetLastError();
}

int phnms_line_iou_f32(const float *pred, const float *target, int64_t num_pred, int64_t num_target, int n_off, float img_w,
                       float length, int aligned, float *out, void *stream) {
    if (num_pred < 0 || num_target < 0 || n_off < 1 || n_off > 250 || (aligned && num_pred != num_target))
        return PHNMS_ERR_BAD_ARG;
    if (num_pred == 0 || num_target == 0) return PHNMS_OK;
    if (!pred || !target || !out) return PHNMS_ERR_BAD_ARG;
    const size_t smem = (size_t)(128 * (n_off + 1) + 32 * n_off) * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(phnms_line_iou_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    phnms_line_iou_kernel<<<(unsigned)((num_pred + 127) / 128), 128, smem, (cudaStream_t)stream>>>(
        pred, target, (int)num_pred, (int)num_target, n_off, img_w, length, aligned, out);
    return (int)cudaGetLastError();
}

int phnms_dynamic_k_assign_f32(const float *cost, const float *iou, int64_t B, int64_t num_priors, int64_t num_gt, int n_candidate_k,
                               int min_k, int binarize, float binarize_at, int64_t *prior_idx, int64_t *gt_idx, int64_t *count,
                               void *stream) {
    if (B < 0 || num_gt < 0 || num_gt > kAssignMaxGt || n_candidate_k < 1 || n_candidate_k > kAssignMaxCand || min_k < 0 ||
        num_priors < n_candidate_k || num_priors > kAssignMaxPriors || B > 0x7fffffff)
        return PHNMS_ERR_BAD_ARG;
    if (B == 0) return PHNMS_OK;
    if (!count || !prior_idx || !gt_idx) return PHNMS_ERR_BAD_ARG;
    if (num_gt == 0) return (int)cudaMemsetAsync(count, 0, (size_t)B * 8, (cudaStream_t)stream);
    if (!cost || !iou) return PHNMS_ERR_BAD_ARG;
    const int threads = (int)((num_priors + 31) / 32) * 32;
    phnms_dynamic_k_assign_kernel<<<(unsigned)B, threads < 64 ? 64 : threads, (size_t)num_gt * sizeof(int), (cudaStream_t)stream>>>(
        cost, iou, (int)num_priors, (int)num_gt, n_candidate_k, min_k, binarize, binarize_at,
        reinterpret_cast<long long *>(prior_idx), reinterpret_cast<long long *>(gt_idx), reinterpret_cast<long long *>(count));
    return (int)cudaGetLastError();
}

int phnms_decode_lanes_f32(const float *rows, const int64_t *num, int64_t T, int64_t K, int n_off, int hdr,
                           const double *prior_ys, double ori_img_h, double cut_height, double *points, int32_t *npoints,
                           float *meta, void *stream) {
    if (T < 0 || K < 0 || (hdr != 6 && hdr != 7) || n_off < 2 || n_off > 96) return PHNMS_ERR_BAD_ARG;
    if (T == 0 || K == 0) return PHNMS_OK;
    if (!rows || !num || !prior_ys || !points || !npoints || !meta || !(ori_img_h > 0.0)) return PHNMS_ERR_BAD_ARG;
    const long long total = (long long)T * K;
    phnms_decode_kernel<<<(unsigned)((total + 3) / 4), 128, 0, (cudaStream_t)stream>>>(
        rows, reinterpret_cast<const long long *>(num), (int)K, hdr, n_off, prior_ys, ori_img_h, cut_height, points, npoints,
        meta, total);
    return (int)cudaGetLastError();
}

// ---- peer-memory plumbing (CUDA IPC) and cross-GPU completion flags -----------------------------------------------------
int phnms_peer_alloc(size_t bytes, void **ptr, unsigned char *handle) {
    if (!ptr || !handle || bytes == 0) return PHNMS_ERR_BAD_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == PHNMS_IPC_HANDLE_BYTES, "IPC handle size");
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);   // a dedicated allocation: IPC handles name whole cudaMalloc allocations
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return (int)e;
    }
    memcpy(handle, &h, sizeof(h));
    *ptr = p;
    return PHNMS_OK;
}

int phnms_peer_open(const unsigned char *handle, void **ptr) {
    if (!ptr || !handle) return PHNMS_ERR_BAD_ARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    return (int)cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
}

int phnms_peer_close(void *ptr) { return ptr ? (int)cudaIpcCloseMemHandle(ptr) : PHNMS_ERR_BAD_ARG; }

int phnms_peer_free(void *ptr) { return ptr ? (int)cudaFree(ptr) : PHNMS_ERR_BAD_ARG; }

int phnms_peer_sync(uint64_t *const *signal_dst, const uint64_t *wait_src, int n, uint64_t signal_epoch,
                    uint64_t wait_epoch, uint64_t timeout_ns, int *status, void *stream) {
    if (n < 1 || n > PHNMS_MAX_DST) return PHNMS_ERR_BAD_ARG;
    if ((signal_epoch && !signal_dst) || (wait_epoch && !wait_src)) return PHNMS_ERR_BAD_ARG;
    if (!signal_epoch && !wait_epoch) return PHNMS_OK;
    PeerSyncArgs a;
    a.n = n;
    a.signal_epoch = signal_epoch;
    a.wait_epoch = wait_epoch;
    a.timeout_ns = timeout_ns ? timeout_ns : 10000000000ull;
    for (int d = 0; d < kMaxCollectDst; ++d)
        a.signal_dst[d] = (signal_epoch && d < n) ? reinterpret_cast<unsigned long long *>(signal_dst[d]) : nullptr;
    a.wait_src = reinterpret_cast<const unsigned long long *>(wait_src);
    a.status = status;
    phnms_peer_sync_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(a);
    return (int)cudaGetLastError();
}

}  // extern "C"

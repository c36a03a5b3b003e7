/*
 * phnms.h -- C ABI of the B200 (sm_100a) lane-NMS library `libphnms.so`.
 *
 * This is the drop-in boundary for PHNet's `libs/ops` lane NMS.  Every entry point takes plain
 * pointers and sizes (no torch types); device pointers are raw CUDA device addresses and `stream`
 * is a `cudaStream_t` passed as `void*`.  The compute entry points never allocate or free device
 * memory and never synchronise the stream: all buffers belong to the caller (PyTorch in the Python
 * mirror, phnet_b200/ops/nms.py).  The only allocating calls are the explicit phnms_peer_* helpers.
 *
 * Reference interfaces replaced (paths relative to the PHNet repository):
 *   libs/ops/nms.py:32-33          nms(boxes, scores, overlap, top_k)
 *   libs/ops/csrc/nms.cpp:44-61    nms_forward(boxes, scores, thresh, top_k)   [pybind module nms_impl]
 *   libs/ops/csrc/nms_kernel.cu:147-192  nms_cuda_forward (launcher), :50-96 nms_kernel, :99-143 nms_collect
 *   libs/ops/csrc/nms.cpp:51       scores.sort(0, true)  (the ordering; see `sort_model`)
 *
 * All functions return PHNMS_OK (0), a negative PHNMS_ERR_* code, or a positive `cudaError_t`.
 * They never throw and keep no global state (re-entrant; any host thread, any stream).
 */
#ifndef PHNMS_H_
#define PHNMS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PHNMS_ABI_VERSION 5

#define PHNMS_OK 0
#define PHNMS_ERR_BAD_ARG (-1)      /* null pointer, negative size, misaligned pointer                         */
#define PHNMS_ERR_N_OFFSETS (-2)    /* n_off outside [1, 250]   (reference: "Wrong number of offsets", nms_kernel.cu:154) */
#define PHNMS_ERR_TOO_MANY (-3)     /* ceil(N/64) >= 1000       (reference: MAX_COL_BLOCKS assert, nms_kernel.cu:158)     */
#define PHNMS_ERR_WORKSPACE (-4)    /* workspace missing or smaller than phnms_workspace_bytes()                */
#define PHNMS_ERR_DEVICE (-5)       /* current device is not compute capability 10.x                           */
#define PHNMS_ERR_TUNING (-6)       /* tuning override cannot be honoured (does not fit shared memory)          */

/* sort_model: how ties between equal scores are ordered (libs/ops/csrc/nms.cpp:51 delegates this to
 * ATen's CUDA sort, which is not a stable sort below 33 elements). */
#define PHNMS_SORT_TORCH_CUDA 0     /* bit-for-bit what torch 2.11 `scores.sort(0, True)` does on CUDA (measured on B200,
                                       tests/golden/torch_cuda_sort.npz): n<=32 ATen's unstable bitonic network,
                                       n>32 stable radix bit order (+NaN first, -NaN last, -0.0 == +0.0)               */
#define PHNMS_SORT_STABLE 1         /* stable descending, NaN of either sign first (torch CPU / numpy semantics)       */
#define PHNMS_SORT_STABLE_RADIX 2   /* stable radix bit order for every n == torch.sort(stable=True) on CUDA          */

/* which device algorithm runs */
#define PHNMS_PATH_AUTO 0
#define PHNMS_PATH_FUSED 1          /* one cluster of CTAs per frame, proposals resident in shared memory, lazy rows   */
#define PHNMS_PATH_TILED 2          /* three kernels: radix order -> 64x64 tile bitmask -> warp-ballot greedy scan     */

/* variants of the fused path */
#define PHNMS_FUSED_SMEM 1          /* cluster per frame, proposals stay in shared memory; any n_off in [1, 250]       */
#define PHNMS_FUSED_REG 2           /* cluster per frame, proposals held in registers, next frame's TMA load overlaps
                                       compute; n_off 36/72, any top_k                                                 */
#define PHNMS_FUSED_STREAM 3        /* the default for n_off 36/72 and 1 <= top_k <= 8: a select kernel runs the greedy scan
                                       on the few best-ranked proposals it needs, a streaming kernel (independent warps,
                                       TMA-fed, no cluster) evaluates every proposal against the kept lanes; frames the
                                       select kernel could not finish are redone by the PHNMS_FUSED_REG kernel          */

#define PHNMS_FUSED_SMALL 4         /* calls of <= 2048 proposals in total with <= 512 per frame (PHNet's own call: one frame of
                                       <= 240 priors): ONE launch, one CTA per frame, rows in registers, one greedy round per
                                       kept lane, any top_k, no workspace                                               */

/* how frames are handed to the persistent clusters of the register-resident kernel */
#define PHNMS_SCHED_STATIC 1        /* cluster c takes frames c, c + n_clusters, ...: fastest when the GPU is not shared         */
#define PHNMS_SCHED_DYNAMIC 2       /* clusters claim frames from a counter: no second wave when another kernel holds some SMs   */

typedef struct phnms_tuning {
    int path;            /* PHNMS_PATH_*                                             (0 = auto) */
    int cluster;         /* CTAs per frame for the fused path: 1,2,4,8,16            (0 = auto) */
    int threads;         /* threads per CTA for the fused path: multiple of 32, <=512 (0 = auto) */
    int max_clusters;    /* cap on resident clusters (persistent grid size)          (0 = auto) */
    int variant;         /* fused path: PHNMS_FUSED_SMEM / _REG / _STREAM / _SMALL   (0 = auto) */
    int schedule;        /* register-resident kernel: PHNMS_SCHED_STATIC / _DYNAMIC   (0 = auto) */
    int stream_warps;    /* streaming kernel: warps per CTA, 1..16                   (0 = auto) */
    int select_cap;      /* select kernel: proposals drawn per frame before the frame is handed to the resume
                            pass, >= 8                                               (0 = auto: 64) */
    int lanes_per_pass;  /* streaming kernel: kept lanes evaluated per pass over the registers, 1/2/4 (0 = auto) */
} phnms_tuning;
/* cluster, threads and schedule describe the cluster kernels; setting any of them (or variant 1/2) selects those kernels. */

typedef struct phnms_plan {
    int path;            /* PHNMS_PATH_FUSED or PHNMS_PATH_TILED */
    int cluster;         /* CTAs per frame (fused) */
    int threads;         /* threads per CTA */
    int rows_per_cta;    /* proposals resident per CTA (fused) */
    int smem_bytes;      /* dynamic shared memory per CTA */
    int grid;            /* CTAs launched */
    int launches;        /* kernel launches one phnms_forward_f32 call makes */
    int variant;         /* PHNMS_FUSED_SMEM / _REG / _STREAM / _SMALL (fused path), 0 otherwise */
    int cols_per_thread; /* proposals held per thread (register-resident variant) */
    int max_active_clusters; /* cudaOccupancyMaxActiveClusters for this launch (0 when no device was queried) */
    size_t workspace_bytes;
} phnms_plan;

int phnms_abi_version(void);
const char *phnms_error_string(int code);

/* Bytes of device workspace `phnms_forward_f32` needs for this shape, whatever top_k is passed: the per-frame kept-lane
 * blocks of the streaming path / candidate blocks of the register-resident cluster kernel (a few KB per frame), the order
 * and bitmask of the tiled path; 0 only for the shared-memory cluster kernel (n_off other than 36 / 72). */
size_t phnms_workspace_bytes(int64_t F, int64_t N, int n_off, const phnms_tuning *tuning /* nullable */);

/* Fills `plan` with what phnms_forward_f32 would launch for this shape on the current device (top_k: as in the call;
 * phnms_plan_query assumes a top_k in [1, 8], PHNet's 4 / 8). */
int phnms_plan_query(int64_t F, int64_t N, int n_off, const phnms_tuning *tuning /* nullable */, phnms_plan *plan);
int phnms_plan_query_topk(int64_t F, int64_t N, int n_off, int64_t top_k, const phnms_tuning *tuning /* nullable */,
                          phnms_plan *plan);

/*
 * Lane NMS over a batch of F independent frames (F = 1 is exactly one reference `nms` call).
 *
 *   props    [F, N, 5+n_off] fp32, contiguous, device.  Row = (logit0, logit1, start_y, start_x, length, x_0..x_{n_off-1})
 *                                                        as produced by get_lanes (libs/models/Router4OL.py:454-458)
 *   scores   [F, N] fp32, contiguous, device
 *   n_valid  [F] int32 device, nullable: frame f uses only its first n_valid[f] rows (NULL = all N)
 *   thresh   `overlap` of the reference call (pixels of mean |dx|); top_k as in the reference (0 = never stop early)
 *   keep     [F, N] int64 device: kept ORIGINAL proposal indices in score order, zero padded        (nms_kernel.cu:118,139-140)
 *   num_keep [F]    int64 device: min(top_k, number kept)                                         (nms_kernel.cu:142)
 *   parent   [F, N] int64 device: 1-based slot of the last kept lane covering each proposal, 0 = none (nms_kernel.cu:123-129)
 *   ws       device workspace of at least phnms_workspace_bytes() bytes (may be NULL when that is 0)
 *   stream   cudaStream_t on which everything is enqueued; the call is asynchronous
 */
int phnms_forward_f32(const float *props, const float *scores, const int32_t *n_valid, int64_t F, int64_t N,
                      int n_off, float thresh, int64_t top_k, int sort_model, int64_t *keep, int64_t *num_keep,
                      int64_t *parent, void *ws, size_t ws_bytes, const phnms_tuning *tuning /* nullable */,
                      void *stream);

/* Same call with a profiling hook: when `trace` (device, int64[trace_len]) is not NULL, thread 0 of CTA 0 of the
 * register-resident fused kernel appends (phase tag, clock64()) pairs at its phase boundaries. */
int phnms_forward_f32_trace(const float *props, const float *scores, const int32_t *n_valid, int64_t F, int64_t N,
                            int n_off, float thresh, int64_t top_k, int sort_model, int64_t *keep, int64_t *num_keep,
                            int64_t *parent, void *ws, size_t ws_bytes, const phnms_tuning *tuning, void *stream,
                            int64_t *trace, int trace_len);

/*
 * Double precision boxes.  The reference instantiates its kernels for double as well (AT_DISPATCH_FLOATING_TYPES,
 * libs/ops/csrc/nms_kernel.cu:171); PHNet itself never passes doubles, so this is a compatibility path (bitmask + scan,
 * not the fused kernels).  The ordering is the caller's: `order` [F, N] int64 is what `scores.sort(0, True)` returned
 * (libs/ops/csrc/nms.cpp:51) -- the Python mirror calls torch for it exactly like the reference does, so ties fall as they
 * do there by construction.  props [F, N, 5 + n_off] fp64; the other arguments as in phnms_forward_f32.
 */
size_t phnms_ordered_f64_workspace_bytes(int64_t F, int64_t N);
int phnms_forward_ordered_f64(const double *props, const int64_t *order, const int32_t *n_valid /* nullable */, int64_t F,
                              int64_t N, int n_off, float thresh, int64_t top_k, int64_t *keep, int64_t *num_keep,
                              int64_t *parent, void *ws, size_t ws_bytes, void *stream);

/*
 * The training-side line IoU: libs/utils/dynamic_assign.py:5-36
 * `line_iou(pred, target, img_w, length=15, aligned)`.  pred [num_pred, n_off], target [num_target, n_off] fp32 device, x in
 * pixels.  aligned != 0: num_pred == num_target, out [num_pred] = IoU of pair i (the LIoU loss term); aligned == 0:
 * out [num_pred, num_target] = the pairwise matrix the dynamic-k assignment consumes (:83-125).  fp32, within 1e-5 relative of
 * the reference (torch's reduction order over the offsets is not sequential).
 */
int phnms_line_iou_f32(const float *pred, const float *target, int64_t num_pred, int64_t num_target, int n_off, float img_w,
                       float length, int aligned, float *out, void *stream);

/*
 * The dynamic-k assignment that consumes that matrix: libs/utils/dynamic_assign.py:83-125
 * `dynamic_k_assign(cost, pair_wise_ious)`; with n_candidate_k / min_k it is also libs/utils/dynamic_assignV2.py:372-405
 * (max_topk / min_topk), and with binarize != 0 `dynamic_k_assign_CF` (dynamic_assign.py:327-370: IoUs >= binarize_at -> 1, else 0;
 * the reference uses n_candidate_k 1, min_k 0 there).  cost, iou [B, num_priors, num_gt] fp32 device (B images at once; the
 * reference takes one).  Writes, per image, count[b] matched priors: prior_idx[b, 0..count) ascending and gt_idx[b, 0..count)
 * (both [B, num_priors] int64, entries beyond count untouched).  num_priors in [n_candidate_k, 1024], num_gt <= 1024,
 * n_candidate_k in [1, 8].  Ties between equal costs go to the lowest index (torch.topk leaves them open); no NaNs.
 */
int phnms_dynamic_k_assign_f32(const float *cost, const float *iou, int64_t B, int64_t num_priors, int64_t num_gt, int n_candidate_k,
                               int min_k, int binarize, float binarize_at, int64_t *prior_idx, int64_t *gt_idx, int64_t *count,
                               void *stream);

/*
 * predictions_to_pred for a whole clip: the tensor part of libs/models/Router4OLV2.py:363-404
 * (hdr == 6) and RouterV4.py:349-392 (hdr == 7) for every kept lane -- start / end rounding, the "extend to the bottom"
 * mask (OpenLane-V models), the -2 fills, selection of the points with x >= 0, the flip, the y rescale (VIL-100 models) --
 * i.e. exactly the `points` array the reference hands to `Lane(points=...)` (libs/utils/lane.py:4-16; the spline inside
 * Lane stays on the host).  One launch, no host sync.
 *   rows     [T, K, hdr + n_off] fp32 device: out_rows of phnms_get_lanes_f32;  num [T] int64: its out_num
 *   prior_ys [n_off] fp64 device: torch.linspace(1, 0, n_off) (fp32, Router4OLV2.py:61) widened to double
 *   ori_img_h, cut_height: arguments of the reference call (used by the VIL-100 variant only, RouterV4.py:378)
 *   points   [T, K, n_off, 2] fp64: (x, y) in `Lane.points` order, zero padded;  npoints [T, K] int32: number of points,
 *            0 where the reference skips the lane (<= 1 point) or the slot is empty;  meta [T, K, 3] fp32: start_x, start_y, conf
 */
int phnms_decode_lanes_f32(const float *rows, const int64_t *num, int64_t T, int64_t K, int n_off, int hdr,
                           const double *prior_ys, double ori_img_h, double cut_height, double *points, int32_t *npoints,
                           float *meta, void *stream);

/*
 * Lane NMS + collection of the kept lanes (the multi-GPU "final collection" of the kept-lane results, done with plain
 * stores over peer memory instead of a collective).  Same as phnms_forward_f32; in addition the compact record of frame f,
 *     int64[top_k + 1] = { keep[f, 0 .. top_k-1] zero padded, num_keep[f] },
 * is stored at row (row0 + f) of each of the n_dst destination buffers ([rows, top_k + 1] int64).  A destination is any
 * address the current device can store to: local memory, or another GPU's buffer mapped into this process
 * (phnms_peer_open) -- then the records travel over NVLink / NVSwitch.  With every rank passing the buffers of all ranks
 * and row0 = rank * F, each rank ends up holding the records of all frames: an all-gather without a collective call.
 * On the streaming, small-frame and register-resident paths the NMS kernels store the records themselves (the select step
 * knows keep[0 .. top_k) first; a frame the resume pass redoes is stored again); the other paths append one small record
 * launch.  Requires top_k >= 1; PHNMS_ERR_BAD_ARG unless width == top_k + 1 and
 * row0 + F <= rows (a record is never stored outside a destination).  Completion across GPUs: phnms_peer_sync.
 */
#define PHNMS_MAX_DST 16
typedef struct phnms_collect {
    int n_dst;                        /* destinations, 1 .. PHNMS_MAX_DST                              */
    int width;                        /* row width of every destination: must equal top_k + 1          */
    int64_t row0;                     /* row of this call's frame 0 in every destination               */
    int64_t rows;                     /* rows of every destination: row0 + F must not exceed it        */
    int64_t *dst[PHNMS_MAX_DST];      /* device-accessible [rows, width] int64 buffers, 8-byte aligned   */
    /* Optional completion across GPUs in the SAME call (see phnms_peer_sync for the flag protocol): when signal_epoch or
     * wait_epoch is non-zero, signal_epoch is stored to signal_dst[d] (d < n_dst: this rank's slot in rank d's flag array;
     * system-scope release, after all record stores of the call) and wait_src[r] >= wait_epoch is awaited for all r < n_dst --
     * by one single-block launch after the NMS kernels when they stored the records themselves, else by the last block of the
     * record kernel.  sync_counter: one device uint32, zero before the first call (left zero).  All zero / NULL: records only. */
    uint64_t signal_epoch, wait_epoch, timeout_ns;
    uint64_t *signal_dst[PHNMS_MAX_DST];
    const uint64_t *wait_src;
    int *status;
    uint32_t *sync_counter;
} phnms_collect;

int phnms_forward_collect_f32(const float *props, const float *scores, const int32_t *n_valid, int64_t F, int64_t N,
                              int n_off, float thresh, int64_t top_k, int sort_model, int64_t *keep, int64_t *num_keep,
                              int64_t *parent, void *ws, size_t ws_bytes, const phnms_tuning *tuning /* nullable */,
                              void *stream, const phnms_collect *collect);

/*
 * Peer-memory plumbing for the collection above (one process per GPU, one node).  These are the only entry points that
 * allocate: phnms_peer_alloc makes a dedicated, zero-filled cudaMalloc allocation on the current device and returns its
 * CUDA IPC handle (64 bytes, to be sent to the other ranks by any host channel); phnms_peer_open maps another rank's
 * allocation into this process (enabling peer access), phnms_peer_close unmaps it, phnms_peer_free releases one's own.
 */
#define PHNMS_IPC_HANDLE_BYTES 64
int phnms_peer_alloc(size_t bytes, void **ptr, unsigned char *handle /* [PHNMS_IPC_HANDLE_BYTES] out */);
int phnms_peer_open(const unsigned char *handle /* [PHNMS_IPC_HANDLE_BYTES] */, void **ptr);
int phnms_peer_close(void *ptr);
int phnms_peer_free(void *ptr);

/*
 * Cross-GPU completion flags, enqueued on `stream` (one tiny kernel, asynchronous).  Every rank owns an array of n
 * uint64 flags in peer-mapped memory.
 *   signal_epoch != 0: store it (system-scope release) to signal_dst[d] for d < n -- this rank's slot in rank d's array.
 *                      Ordered after everything earlier work of this stream stored to those peers.
 *   wait_epoch   != 0: spin until wait_src[r] >= wait_epoch for all r < n (system-scope acquire) -- every rank has
 *                      signalled that epoch -- or until timeout_ns (0 = 10 s) has passed; on timeout *status (device int,
 *                      nullable) is set to 1 + the index of the slot that did not arrive and the kernel returns.
 * Epochs must increase from call to call (flags are never reset).
 */
int phnms_peer_sync(uint64_t *const *signal_dst /* host array [n] of device pointers */, const uint64_t *wait_src, int n,
                    uint64_t signal_epoch, uint64_t wait_epoch, uint64_t timeout_ns, int *status, void *stream);

/*
 * The ordering alone (libs/ops/csrc/nms.cpp:51): order[f, i] = original index of the i-th proposal of frame f
 * in descending score order under `sort_model`; entries beyond n_valid[f] are zero.
 * `ws` must hold phnms_order_workspace_bytes(F, N) bytes.
 */
size_t phnms_order_workspace_bytes(int64_t F, int64_t N);
int phnms_order_f32(const float *scores, const int32_t *n_valid, int64_t F, int64_t N, int sort_model,
                    int64_t *order, void *ws, size_t ws_bytes, void *stream);

/*
 * get_lanes for a whole clip (the callers either side of the op): libs/models/Router4OL.py:447-470 and
 * RouterV4.py:404-428 for T frames in 4 launches and no host sync.
 *
 *   pred       [T, A, hdr + n_off] fp32 device: raw head output per prior: (logit0, logit1, start_y, start_x, theta, length,
 *              [invalid_length when hdr == 7], x_0 .. x_{n_off-1}), x and start_x normalised to [0, 1]
 *   hdr        6 (OpenLane-V models) or 7 (VIL-100 models)
 *   score = softmax(logits)[1]; priors with score >= conf_threshold survive, in prior order; the NMS rows are
 *   (logit0, logit1, start_y, start_x * (img_w-1), length * (n_off-1), x * (img_w-1)); overlap = nms_thres; top_k = max_lanes
 *   out_rows   [T, top_k, hdr + n_off]: pred[kept] with columns 5 .. hdr-1 replaced by round(col * (n_off-1)), zero padded
 *   out_num    [T] int64 lanes kept;  out_index [T, top_k] int64 kept priors' indices in the unfiltered frame
 *   keep_inds  [T, A] uint8: the confidence mask (`keep_inds` of the reference)
 */
size_t phnms_get_lanes_workspace_bytes(int64_t T, int64_t A, int n_off, const phnms_tuning *tuning /* nullable */);
int phnms_get_lanes_f32(const float *pred, int64_t T, int64_t A, int n_off, int hdr, float conf_threshold, float img_w,
                        float nms_thres, int64_t top_k, int sort_model, float *out_rows, int64_t *out_num,
                        int64_t *out_index, unsigned char *keep_inds, void *ws, size_t ws_bytes,
                        const phnms_tuning *tuning /* nullable */, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PHNMS_H_ */

/*
 * oracle/lane_nms_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * CPU restatement (plain C, scalar) of PHNet's lane-NMS op `libs/ops`:
 *     nms(boxes, scores, overlap, top_k) -> (keep, num_to_keep, parent_object_index)
 * Citations are relative to /root/reference/.
 *
 *   ordering        libs/ops/csrc/nms.cpp:51          scores.sort(0, true)  (ATen CUDA sort, un-vendored;
 *                                                    modelled after torch 2.11 ATen/native/cuda/SortUtils.cuh)
 *   pair predicate  libs/ops/csrc/nms_kernel.cu:26-48 devIoU  (mean |dx| over the shared y-range < threshold)
 *   tile mask       libs/ops/csrc/nms_kernel.cu:50-96 nms_kernel (strict upper triangle in SORTED order)
 *   greedy scan     libs/ops/csrc/nms_kernel.cu:99-143 nms_collect (keep / parent / zero padding / min(top_k, n))
 *
 * The only things allowed to call this file are tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs.  The product path
 * (phnet_b200/csrc) never links or loads it.
 *
 * Parity pin: the reference ships no tests or golden vectors for this op
 * (SURVEY.md section 4 / 8c).  This restatement is pinned against the reference's own
 * CUDA kernels compiled for sm_100a (oracle/build_ref.py -> oracle/_ref/) and run
 * on a B200; the outputs are committed under tests/golden/ (see
 * tests/golden/make_ref_fixtures.py).
 *
 * GPU arithmetic that has to be emulated on x86 (from the PTX of the reference build):
 *   (int)(double)      -> F2I.F64.TRUNC   : truncates toward zero, SATURATES, NaN -> 0x80000000 (INT_MIN).
 *                         (PTX documents NaN -> 0 for cvt; the B200 measurably returns INT_MIN for the f64 source:
 *                          the live reference op only matches this restatement with INT_MIN, see tests/golden.)
 *   int add/sub        -> wraps (two's complement)
 *   unsigned char i    -> (5 + start) & 255
 *   float accumulation -> sequential fp32 adds in ascending offset order, no FMA
 * Build with -ffp-contract=off and without -ffast-math (oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
#include <unistd.h>

#define THREADS_PER_BLOCK 64 /* nms_kernel.cu:18  sizeof(unsigned long long) * 8 */
#define MAX_COL_BLOCKS 1000  /* nms_kernel.cu:10 */

/* ---- GPU integer conversion: cvt.rzi.s32.f64 -------------------------------------- */
static int32_t cvt_rzi_s32_f64(double d) {
    if (d != d) return INT32_MIN; /* measured on B200: F2I.F64.TRUNC(NaN) = 0x80000000 (tests/golden) */
    if (d >= 2147483647.0) return INT32_MAX;
    if (d <= -2147483648.0) return INT32_MIN;
    return (int32_t)d;
}

static int32_t wrap_add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static int32_t wrap_sub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }

/* nms_kernel.cu:29-30   (int)(a[2] * N_STRIPS - DATASET_OFFSET + 0.5)
 * fp32 multiply, then fp64 add of 0.5, then truncation. */
int32_t phoracle_lane_start(const float *row, int n_off) {
    float m = row[2] * (float)(n_off - 1);
    return cvt_rzi_s32_f64((double)m + 0.5);
}

/* nms_kernel.cu:32-33   start + a[4] - 1 + 0.5 - ((a[4] - 1) < 0)
 * int->fp32, fp32 add, fp32 add of -1, then fp64 +0.5, fp64 -{0,1}, truncation. */
int32_t phoracle_lane_end(const float *row, int32_t start) {
    float s = (float)start + row[4];
    float t = s - 1.0f;
    float lm1 = row[4] - 1.0f;
    double e = (double)t + 0.5;
    e = e - (double)((lm1 < 0.0f) ? 1 : 0);
    return cvt_rzi_s32_f64(e);
}

/* devIoU with the per-lane start/end (pure functions of one lane) hoisted out of the pair loop. */
static int pred_pre(const float *a, const float *b, int32_t start_a, int32_t end_a, int32_t start_b,
                    int32_t end_b, int n_off, float threshold) {
    const int32_t start = start_a > start_b ? start_a : start_b;            /* :31 */
    int32_t end = end_a < end_b ? end_a : end_b;                             /* :34 */
    if (end > n_off - 1) end = n_off - 1;
    if (end < start) return 0;                                               /* :36 */
    float dist = 0.0f;
    /* :38  for (unsigned char i = 5 + start; i <= 5 + end; ++i)
     * end <= n_off-1 <= 249 so the counter never wraps INSIDE the loop. */
    uint8_t i = (uint8_t)((uint32_t)start + 5u);
    const int32_t last = wrap_add(end, 5);
    for (; (int32_t)i <= last; ++i) {
        if (a[i] < b[i]) {
            dist += b[i] - a[i];
        } else {
            dist += a[i] - b[i];
        }
    }
    const int32_t len = wrap_add(wrap_sub(end, start), 1);                   /* :46 */
    const float lim = threshold * (float)len;
    return dist < lim;
}

/* devIoU, nms_kernel.cu:26-48.  a = the higher-ranked lane (row), b = the column lane. */
int phoracle_pred(const float *a, const float *b, int n_off, float threshold) {
    const int32_t start_a = phoracle_lane_start(a, n_off);
    const int32_t start_b = phoracle_lane_start(b, n_off);
    return pred_pre(a, b, start_a, phoracle_lane_end(a, start_a), start_b,
                    phoracle_lane_end(b, start_b), n_off, threshold);
}

static void lane_bounds(const float *props, int64_t n, int n_off, int32_t *se) {
    for (int64_t r = 0; r < n; ++r) {
        const float *row = props + r * (5 + n_off);
        se[2 * r] = phoracle_lane_start(row, n_off);
        se[2 * r + 1] = phoracle_lane_end(row, se[2 * r]);
    }
}

/* ---- ordering: model of at::Tensor::sort(0, descending=true) on CUDA ------------------ */
static uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

/* ascending-radix twiddle (cub Traits<float>::TwiddleIn), with -0.0 -> +0.0
 * (cub BaseDigitExtractor::ProcessFloatMinusZero) */
static uint32_t twiddle_asc(uint32_t u) {
    if (u == 0x80000000u) u = 0u;
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

/* mode 0: comparator order, GTOp<float,true> (SortingCommon.cuh:47-53): NaN (any sign) first, NaNs tie
 * mode 1: radix order (cub BlockRadixSort::SortDescending): by bit pattern, +NaN first, -NaN last   */
uint32_t phoracle_key_desc(float s, int mode) {
    uint32_t u = f2u(s);
    if (mode == 0 && s != s) return 0u;
    return ~twiddle_asc(u);
}

static int gt_nan(float l, float r) { /* GTOp<float, true> */
    return ((l != l) && !(r != r)) || (l > r);
}

static int cmp_u64(const void *x, const void *y) {
    uint64_t a = *(const uint64_t *)x, b = *(const uint64_t *)y;
    return (a > b) - (a < b);
}

/* sort_model: 0 = torch 2.11 CUDA `sort(0, True)`: n<=32 unstable bitonic network, n>32 stable radix bit order
 *                 (pinned by tests/golden/torch_cuda_sort.npz, generated on a B200)
 *             1 = stable, comparator order: NaN of either sign first (torch CPU / numpy semantics)
 *             2 = stable radix bit order for every n (== torch.sort(stable=True) on CUDA): +NaN first, -NaN last */
void phoracle_order(const float *scores, int64_t n, int64_t *order, int sort_model) {
    if (n <= 0) return;
    if (n == 1) { order[0] = 0; return; }
    if (sort_model == 0 && n <= 32) {
        /* bitonicSortKVInPlace<block_dim_x = 16>, SortUtils.cuh:45-163; Power2SortSize = 32 */
        float keys[32]; int64_t vals[32]; int valid[32];
        for (int i = 0; i < 32; ++i) {
            valid[i] = i < n;
            keys[i] = valid[i] ? scores[i] : 0.0f;
            vals[i] = valid[i] ? i : 0;
        }
#define BSWAP(pa, pb, dir)                                                           \
    do {                                                                             \
        int sw = (gt_nan(keys[pa], keys[pb]) && valid[pa]) || !valid[pb];            \
        if (sw == (dir)) {                                                           \
            float tk = keys[pa]; keys[pa] = keys[pb]; keys[pb] = tk;                 \
            int64_t tv = vals[pa]; vals[pa] = vals[pb]; vals[pb] = tv;               \
            int tb = valid[pa]; valid[pa] = valid[pb]; valid[pb] = tb;               \
        }                                                                            \
    } while (0)
        for (unsigned size = 2; size < 32; size *= 2) {
            for (unsigned stride = size / 2; stride > 0; stride /= 2) {
                for (unsigned tx = 0; tx < 16; ++tx) {
                    int flag = (tx & (size / 2)) != 0;
                    unsigned pos = 2 * tx - (tx & (stride - 1));
                    BSWAP(pos, pos + stride, flag);
                }
            }
        }
        for (unsigned stride = 16; stride > 0; stride /= 2) {
            for (unsigned tx = 0; tx < 16; ++tx) {
                unsigned pos = 2 * tx - (tx & (stride - 1));
                BSWAP(pos, pos + stride, 0);
            }
        }
#undef BSWAP
        for (int64_t i = 0; i < n; ++i) order[i] = vals[i];
        return;
    }
    int mode = (sort_model == 1) ? 0 : 1;
    uint64_t *tmp = (uint64_t *)malloc((size_t)n * sizeof(uint64_t));
    for (int64_t i = 0; i < n; ++i)
        tmp[i] = ((uint64_t)phoracle_key_desc(scores[i], mode) << 32) | (uint64_t)(uint32_t)i;
    qsort(tmp, (size_t)n, sizeof(uint64_t), cmp_u64);
    for (int64_t i = 0; i < n; ++i) order[i] = (int64_t)(tmp[i] & 0xffffffffu);
    free(tmp);
}

/* ---- literal restatement: full upper-triangle bitmask, then the serial collect ---------- */
/* returns 0, or -1 bad argument (mirrors the TORCH_CHECK / AT_ASSERTM of nms_kernel.cu:154,158) */
int phoracle_nms_literal(const float *props, const int64_t *idx, int64_t n, int n_off, float thr,
                         int64_t top_k, int64_t *keep, int64_t *num_to_keep, int64_t *parent) {
    const int64_t P = 5 + n_off;
    const int64_t col_blocks = (n + THREADS_PER_BLOCK - 1) / THREADS_PER_BLOCK;   /* :156 */
    if (n_off < 1 || n_off > 250) return -1;
    if (col_blocks >= MAX_COL_BLOCKS) return -1;                                   /* :158 */
    if (n == 0) { *num_to_keep = 0; return 0; }
    uint64_t *mask = (uint64_t *)malloc((size_t)(n * col_blocks) * sizeof(uint64_t));
    int32_t *se = (int32_t *)malloc((size_t)n * 2 * sizeof(int32_t));
    lane_bounds(props, n, n_off, se);
    /* nms_kernel: grid (col_blocks, col_blocks), 64 threads */
    for (int64_t row_start = 0; row_start < col_blocks; ++row_start) {
        for (int64_t col_start = row_start; col_start < col_blocks; ++col_start) { /* :56 */
            int64_t row_size = n - row_start * 64; if (row_size > 64) row_size = 64;
            int64_t col_size = n - col_start * 64; if (col_size > 64) col_size = 64;
            for (int64_t tx = 0; tx < row_size; ++tx) {
                const int64_t cur = 64 * row_start + tx;
                const int64_t ra = idx[cur];
                const float *a = props + ra * P;                                    /* :81 */
                uint64_t t = 0;
                int64_t s = (row_start == col_start) ? tx + 1 : 0;                  /* :85-87 */
                for (int64_t i = s; i < col_size; ++i) {
                    const int64_t rb = idx[64 * col_start + i];
                    const float *b = props + rb * P;                                /* :66 */
                    if (pred_pre(a, b, se[2 * ra], se[2 * ra + 1], se[2 * rb], se[2 * rb + 1], n_off, thr))
                        t |= 1ULL << i;
                }
                mask[cur * col_blocks + col_start] = t;                             /* :94 */
            }
        }
    }
    /* nms_collect */
    uint64_t *remv = (uint64_t *)calloc((size_t)col_blocks, sizeof(uint64_t));      /* :103-105 */
    int64_t nk = 0;
    for (int64_t i = 0; i < n; ++i) parent[i] = 0;                                  /* :107-109 */
    for (int64_t i = 0; i < n; ++i) {
        int64_t nblock = i / 64, inblock = i % 64;
        if (!(remv[nblock] & (1ULL << inblock))) {                                   /* :116 */
            keep[nk] = idx[i];
            const uint64_t *p = mask + i * col_blocks;
            for (int64_t j = nblock; j < col_blocks; ++j) remv[j] |= p[j];          /* :120-122 */
            for (int64_t j = i; j < n; ++j)                                          /* :123-128 */
                if (p[j / 64] & (1ULL << (j % 64))) parent[idx[j]] = nk + 1;
            parent[idx[i]] = nk + 1;                                                 /* :129 */
            nk++;
            if (nk == top_k) break;                                                  /* :133 */
        }
    }
    for (int64_t i = nk; i < n; ++i) keep[i] = 0;                                    /* :139-140 */
    *num_to_keep = top_k < nk ? top_k : nk;                                          /* :142 */
    free(remv);
    free(mask);
    free(se);
    return 0;
}

/* ---- lazy restatement: only the rows nms_collect actually reads (rows of kept lanes) ------
 * Same outputs as the literal form (asserted in tests/test_oracle.py); used where the literal
 * O(n^2) form would take minutes. */
int phoracle_nms_lazy(const float *props, const int64_t *idx, int64_t n, int n_off, float thr,
                      int64_t top_k, int64_t *keep, int64_t *num_to_keep, int64_t *parent) {
    const int64_t P = 5 + n_off;
    const int64_t col_blocks = (n + THREADS_PER_BLOCK - 1) / THREADS_PER_BLOCK;
    if (n_off < 1 || n_off > 250) return -1;
    if (col_blocks >= MAX_COL_BLOCKS) return -1;
    if (n == 0) { *num_to_keep = 0; return 0; }
    uint8_t *removed = (uint8_t *)calloc((size_t)n, 1);
    int32_t *se = (int32_t *)malloc((size_t)n * 2 * sizeof(int32_t));
    lane_bounds(props, n, n_off, se);
    int64_t nk = 0;
    for (int64_t i = 0; i < n; ++i) parent[i] = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (removed[i]) continue;
        keep[nk] = idx[i];
        const int64_t ra = idx[i];
        const float *a = props + ra * P;
        for (int64_t j = i + 1; j < n; ++j) {
            const int64_t rb = idx[j];
            if (pred_pre(a, props + rb * P, se[2 * ra], se[2 * ra + 1], se[2 * rb], se[2 * rb + 1], n_off, thr)) {
                removed[j] = 1;
                parent[idx[j]] = nk + 1;
            }
        }
        parent[idx[i]] = nk + 1;
        nk++;
        if (nk == top_k) break;
    }
    for (int64_t i = nk; i < n; ++i) keep[i] = 0;
    *num_to_keep = top_k < nk ? top_k : nk;
    free(removed);
    free(se);
    return 0;
}

/* ---- double precision boxes: nms_kernel<double> / devIoU<double> (nms_kernel.cu:26-48,171) ------------------------------
 * Arithmetic as the reference's double instantiation compiles (SASS of oracle/_ref): the start is ONE fused multiply-add
 * (nvcc contracts `a[2] * N_STRIPS + 0.5` under its default -fmad=true: DFMA), the end a chain of fp64 adds, both through
 * F2I.F64.TRUNC; fp64 sequential distance; the limit is the FP32 product threshold * len (devIoU takes `const float
 * threshold`) widened to double.  The ordering `idx` is given (the reference gets it from torch, nms.cpp:51). */
static int32_t lane_start_f64(const double *row, int n_off) {
    return cvt_rzi_s32_f64(fma(row[2], (double)(n_off - 1), 0.5));
}
static int32_t lane_end_f64(const double *row, int32_t start) {
    double e = (((double)start + row[4]) - 1.0) + 0.5;
    e = e - (double)(((row[4] - 1.0) < 0.0) ? 1 : 0);
    return cvt_rzi_s32_f64(e);
}
static int pred_f64(const double *a, const double *b, int32_t sa, int32_t ea, int32_t sb, int32_t eb, int n_off,
                    float threshold) {
    const int32_t start = sa > sb ? sa : sb;
    int32_t end = ea < eb ? ea : eb;
    if (end > n_off - 1) end = n_off - 1;
    if (end < start) return 0;
    double dist = 0.0;
    uint8_t i = (uint8_t)((uint32_t)start + 5u);
    const int32_t last = wrap_add(end, 5);
    for (; (int32_t)i <= last; ++i) {
        if (a[i] < b[i]) dist += b[i] - a[i];
        else dist += a[i] - b[i];
    }
    const float lim = threshold * (float)wrap_add(wrap_sub(end, start), 1);
    return dist < (double)lim;
}
int phoracle_nms_ordered_f64(const double *props, const int64_t *idx, int64_t n, int n_off, float thr, int64_t top_k,
                             int64_t *keep, int64_t *num_to_keep, int64_t *parent) {
    const int64_t P = 5 + n_off;
    if (n_off < 1 || n_off > 250) return -1;
    if ((n + THREADS_PER_BLOCK - 1) / THREADS_PER_BLOCK >= MAX_COL_BLOCKS) return -1;
    if (n == 0) { *num_to_keep = 0; return 0; }
    uint8_t *removed = (uint8_t *)calloc((size_t)n, 1);
    int32_t *se = (int32_t *)malloc((size_t)n * 2 * sizeof(int32_t));
    for (int64_t r = 0; r < n; ++r) {
        se[2 * r] = lane_start_f64(props + r * P, n_off);
        se[2 * r + 1] = lane_end_f64(props + r * P, se[2 * r]);
    }
    int64_t nk = 0;
    for (int64_t i = 0; i < n; ++i) parent[i] = 0;
    for (int64_t i = 0; i < n; ++i) {          /* the rows nms_collect reads (:116-129), see phoracle_nms_lazy */
        if (removed[i]) continue;
        keep[nk] = idx[i];
        const int64_t ra = idx[i];
        for (int64_t j = i + 1; j < n; ++j) {
            const int64_t rb = idx[j];
            if (pred_f64(props + ra * P, props + rb * P, se[2 * ra], se[2 * ra + 1], se[2 * rb], se[2 * rb + 1], n_off, thr)) {
                removed[j] = 1;
                parent[idx[j]] = nk + 1;
            }
        }
        parent[idx[i]] = nk + 1;
        nk++;
        if (nk == top_k) break;
    }
    for (int64_t i = nk; i < n; ++i) keep[i] = 0;
    *num_to_keep = top_k < nk ? top_k : nk;
    free(removed);
    free(se);
    return 0;
}

/* ---- whole op on one frame: sort model + NMS ---------------------------------------------- */
int phoracle_nms(const float *props, const float *scores, int64_t n, int n_off, float thr,
                 int64_t top_k, int sort_model, int lazy, int64_t *keep, int64_t *num_to_keep,
                 int64_t *parent) {
    if (n == 0) { *num_to_keep = 0; return 0; }
    int64_t *idx = (int64_t *)malloc((size_t)n * sizeof(int64_t));
    phoracle_order(scores, n, idx, sort_model);
    int rc = lazy ? phoracle_nms_lazy(props, idx, n, n_off, thr, top_k, keep, num_to_keep, parent)
                  : phoracle_nms_literal(props, idx, n, n_off, thr, top_k, keep, num_to_keep, parent);
    free(idx);
    return rc;
}

/* batch of F frames, each [n_pad, 5+n_off] with n_valid[f] real rows (NULL -> n_pad); frames run in
 * parallel on `threads` host threads (pthreads, dynamic frame claiming).  keep/parent rows are n_pad
 * wide; entries beyond n_valid[f] are zero. */
typedef struct {
    const float *props, *scores;
    const int32_t *n_valid;
    int64_t F, n_pad, top_k;
    int n_off, sort_model, lazy;
    float thr;
    int64_t *keep, *num_to_keep, *parent;
    int64_t next; /* claimed with __atomic_fetch_add */
    int rc;
} batch_job;

static void *batch_worker(void *arg) {
    batch_job *j = (batch_job *)arg;
    const int64_t P = 5 + j->n_off;
    for (;;) {
        int64_t f = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (f >= j->F) break;
        int64_t n = j->n_valid ? j->n_valid[f] : j->n_pad;
        if (n < 0) n = 0;
        if (n > j->n_pad) n = j->n_pad;
        int64_t *k = j->keep + f * j->n_pad, *p = j->parent + f * j->n_pad;
        for (int64_t i = 0; i < j->n_pad; ++i) { k[i] = 0; p[i] = 0; }
        int rc = phoracle_nms(j->props + f * j->n_pad * P, j->scores + f * j->n_pad, n, j->n_off, j->thr,
                              j->top_k, j->sort_model, j->lazy, k, j->num_to_keep + f, p);
        if (rc) __atomic_store_n(&j->rc, -1, __ATOMIC_RELAXED);
    }
    return NULL;
}

int phoracle_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

int phoracle_nms_batched(const float *props, const float *scores, const int32_t *n_valid, int64_t F,
                         int64_t n_pad, int n_off, float thr, int64_t top_k, int sort_model, int lazy,
                         int threads, int64_t *keep, int64_t *num_to_keep, int64_t *parent) {
    batch_job job = {props, scores, n_valid, F, n_pad, top_k, n_off, sort_model, lazy, thr,
                     keep, num_to_keep, parent, 0, 0};
    if (threads <= 0) threads = phoracle_max_threads();
    if (threads > F) threads = (int)(F > 0 ? F : 1);
    if (threads > 1024) threads = 1024;
    pthread_t *tid = (pthread_t *)malloc((size_t)threads * sizeof(pthread_t));
    int started = 0;
    for (int t = 1; t < threads; ++t)
        if (pthread_create(&tid[started], NULL, batch_worker, &job) == 0) started++;
    batch_worker(&job);
    for (int t = 0; t < started; ++t) pthread_join(tid[t], NULL);
    free(tid);
    return job.rc;
}

"""Python face of the CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT.

Wraps oracle/liblane_nms_oracle.so (built by oracle/Makefile from
lane_nms_oracle.c) and adds `nms_py`, an independent pure-Python/numpy
restatement used on small cases to cross-check the C code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this module.  Reference citations
(relative to /root/reference/): libs/ops/nms.py:32-33, libs/ops/csrc/nms.cpp:44-57,
libs/ops/csrc/nms_kernel.cu:26-143.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(HERE, "liblane_nms_oracle.so")

# sort models (see phoracle_order)
SORT_TORCH_CUDA = 0      # what `scores.sort(0, True)` does on CUDA in torch 2.11 (n<=32 bitonic, else stable radix)
SORT_STABLE_CMP = 1      # stable, comparator order (NaN of any sign first; torch CPU semantics)
SORT_STABLE_RADIX = 2    # stable, radix bit order (+NaN first, -NaN last) == torch.sort(stable=True) on CUDA

_f32p = ctypes.POINTER(ctypes.c_float)
_i64p = ctypes.POINTER(ctypes.c_int64)
_i32p = ctypes.POINTER(ctypes.c_int32)


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "lane_nms_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-s", "-B" if force else "-s"], check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.phoracle_pred.argtypes = [_f32p, _f32p, ctypes.c_int, ctypes.c_float]
        L.phoracle_pred.restype = ctypes.c_int
        L.phoracle_lane_start.argtypes = [_f32p, ctypes.c_int]
        L.phoracle_lane_start.restype = ctypes.c_int32
        L.phoracle_lane_end.argtypes = [_f32p, ctypes.c_int32]
        L.phoracle_lane_end.restype = ctypes.c_int32
        L.phoracle_order.argtypes = [_f32p, ctypes.c_int64, _i64p, ctypes.c_int]
        L.phoracle_order.restype = None
        for name in ("phoracle_nms_literal", "phoracle_nms_lazy"):
            fn = getattr(L, name)
            fn.argtypes = [_f32p, _i64p, ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_int64,
                           _i64p, _i64p, _i64p]
            fn.restype = ctypes.c_int
        L.phoracle_nms_batched.argtypes = [_f32p, _f32p, _i32p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                                           ctypes.c_float, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_int, _i64p, _i64p, _i64p]
        L.phoracle_nms_batched.restype = ctypes.c_int
        L.phoracle_max_threads.restype = ctypes.c_int
        L.phoracle_nms_ordered_f64.argtypes = [ctypes.POINTER(ctypes.c_double), _i64p, ctypes.c_int64, ctypes.c_int,
                                               ctypes.c_float, ctypes.c_int64, _i64p, _i64p, _i64p]
        L.phoracle_nms_ordered_f64.restype = ctypes.c_int
        _lib = L
    return _lib


def _f32(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float32))


def _ptr(a, t):
    return a.ctypes.data_as(t)


def nms_f64(props, idx, thr: float, top_k: int):
    """One frame of DOUBLE boxes with a given ordering `idx` (what `scores.sort(0, True)` returned, nms.cpp:51):
    nms_kernel<double> + nms_collect.  Returns (keep[N] i64, num_to_keep int, parent[N] i64)."""
    p = np.ascontiguousarray(np.asarray(props, dtype=np.float64))
    n, P = p.shape
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    keep = np.zeros(n, dtype=np.int64)
    parent = np.zeros(n, dtype=np.int64)
    num = np.zeros(1, dtype=np.int64)
    rc = lib().phoracle_nms_ordered_f64(p.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), _ptr(idx, _i64p), n, P - 5, thr,
                                        top_k, _ptr(keep, _i64p), _ptr(num, _i64p), _ptr(parent, _i64p))
    if rc != 0:
        raise RuntimeError("oracle: bad argument")
    return keep, int(num[0]), parent


def max_threads() -> int:
    return int(lib().phoracle_max_threads())


def pred(a, b, n_off: int, thr: float) -> bool:
    a, b = _f32(a), _f32(b)
    assert a.shape == b.shape == (5 + n_off,)
    return bool(lib().phoracle_pred(_ptr(a, _f32p), _ptr(b, _f32p), n_off, thr))


def lane_bounds(row, n_off: int):
    row = _f32(row)
    s = lib().phoracle_lane_start(_ptr(row, _f32p), n_off)
    e = lib().phoracle_lane_end(_ptr(row, _f32p), s)
    return int(s), int(e)


def order(scores, sort_model: int = SORT_TORCH_CUDA) -> np.ndarray:
    s = _f32(scores).reshape(-1)
    out = np.zeros(s.shape[0], dtype=np.int64)
    lib().phoracle_order(_ptr(s, _f32p), s.shape[0], _ptr(out, _i64p), sort_model)
    return out


def nms(props, scores, thr: float, top_k: int, sort_model: int = SORT_TORCH_CUDA, lazy: bool = False,
        idx=None):
    """One frame.  Returns (keep[N] i64, num_to_keep int, parent[N] i64) like libs/ops/nms.py:32."""
    p = _f32(props)
    n, P = p.shape
    n_off = P - 5
    if idx is None:
        idx = order(scores, sort_model)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    keep = np.zeros(n, dtype=np.int64)
    parent = np.zeros(n, dtype=np.int64)
    num = np.zeros(1, dtype=np.int64)
    fn = lib().phoracle_nms_lazy if lazy else lib().phoracle_nms_literal
    rc = fn(_ptr(p, _f32p), _ptr(idx, _i64p), n, n_off, thr, top_k, _ptr(keep, _i64p), _ptr(num, _i64p),
            _ptr(parent, _i64p))
    if rc != 0:
        raise RuntimeError("oracle: bad argument (n_off outside [1,250] or col_blocks >= 1000)")
    return keep, int(num[0]), parent


def nms_batched(props, scores, n_valid, thr: float, top_k: int, sort_model: int = SORT_TORCH_CUDA,
                lazy: bool = False, threads: int = 0):
    """Batch of frames [F, N, 5+No]; returns (keep[F,N], num[F], parent[F,N])."""
    p = _f32(props)
    F, n, P = p.shape
    s = _f32(scores).reshape(F, n)
    nv = None if n_valid is None else np.ascontiguousarray(n_valid, dtype=np.int32)
    keep = np.zeros((F, n), dtype=np.int64)
    parent = np.zeros((F, n), dtype=np.int64)
    num = np.zeros(F, dtype=np.int64)
    rc = lib().phoracle_nms_batched(_ptr(p, _f32p), _ptr(s, _f32p), None if nv is None else _ptr(nv, _i32p),
                                    F, n, P - 5, thr, top_k, sort_model, int(lazy), threads,
                                    _ptr(keep, _i64p), _ptr(num, _i64p), _ptr(parent, _i64p))
    if rc != 0:
        raise RuntimeError("oracle: bad argument")
    return keep, num, parent


# ---------------------------------------------------------------------------------------------
# Independent pure-Python restatement (small cases only).  Written from the reference source, not
# from the C file above: nms_kernel.cu:26-48 (devIoU), :50-96 (tile mask), :99-143 (collect).
# ---------------------------------------------------------------------------------------------
def _cvt_rzi(d: float) -> int:
    if d != d:
        return -2147483648  # B200 F2I.F64.TRUNC(NaN) = INT_MIN (pinned by tests/golden)
    if d >= 2147483647.0:
        return 2147483647
    if d <= -2147483648.0:
        return -2147483648
    return int(d)  # int() truncates toward zero


def _wrap32(v: int) -> int:
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v & 0x80000000 else v


def _pred_py(a: np.ndarray, b: np.ndarray, n_off: int, thr: np.float32) -> bool:
    f32 = np.float32
    with np.errstate(all="ignore"):
        n_strips = f32(n_off - 1)
        sa = _cvt_rzi(float(f32(a[2] * n_strips)) + 0.5)
        sb = _cvt_rzi(float(f32(b[2] * n_strips)) + 0.5)
        start = max(sa, sb)

        def _end(x, s):
            t = f32(f32(f32(s) + x[4]) - f32(1.0))
            neg = 1.0 if f32(x[4] - f32(1.0)) < f32(0.0) else 0.0
            return _cvt_rzi(float(t) + 0.5 - neg)

        end = min(min(_end(a, sa), _end(b, sb)), n_off - 1)
        if end < start:
            return False
        dist = f32(0.0)
        i = (5 + start) & 255
        last = _wrap32(5 + end)
        while i <= last:
            if a[i] < b[i]:
                dist = f32(dist + f32(b[i] - a[i]))
            else:
                dist = f32(dist + f32(a[i] - b[i]))
            i += 1
        lim = f32(thr * f32(_wrap32(end - start + 1)))
        return bool(dist < lim)


def nms_py(props, scores, thr: float, top_k: int, idx=None):
    """Literal 64x64-tile mask + serial collect, in Python.  O(N^2 * No) interpreter steps."""
    p = _f32(props)
    n, P = p.shape
    n_off = P - 5
    thr = np.float32(thr)
    if idx is None:
        idx = order(scores, SORT_TORCH_CUDA)
    cb = (n + 63) // 64
    mask = [[0] * cb for _ in range(n)]
    for rs in range(cb):
        for cs in range(rs, cb):
            row_size = min(n - rs * 64, 64)
            col_size = min(n - cs * 64, 64)
            for tx in range(row_size):
                cur = 64 * rs + tx
                a = p[idx[cur]]
                t = 0
                for i in range(tx + 1 if rs == cs else 0, col_size):
                    if _pred_py(a, p[idx[64 * cs + i]], n_off, thr):
                        t |= 1 << i
                mask[cur][cs] = t
    remv = [0] * cb
    keep = np.zeros(n, dtype=np.int64)
    parent = np.zeros(n, dtype=np.int64)
    nk = 0
    for i in range(n):
        nb, ib = divmod(i, 64)
        if not (remv[nb] >> ib) & 1:
            keep[nk] = idx[i]
            row = mask[i]
            for j in range(nb, cb):
                remv[j] |= row[j]
            for j in range(i, n):
                if (row[j // 64] >> (j % 64)) & 1:
                    parent[idx[j]] = nk + 1
            parent[idx[i]] = nk + 1
            nk += 1
            if nk == top_k:
                break
    return keep, min(top_k, nk), parent

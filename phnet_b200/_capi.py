"""ctypes binding of libphnms.so (include/phnms.h).  Fails loudly: there is no CPU or PyTorch fallback."""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("PHNMS_SO") or os.path.join(HERE, "csrc", "libphnms.so")   # PHNMS_SO: A/B testing of builds

PATH_AUTO, PATH_FUSED, PATH_TILED = 0, 1, 2
FUSED_SMEM, FUSED_REG, FUSED_STREAM, FUSED_SMALL = 1, 2, 3, 4
SCHED_STATIC, SCHED_DYNAMIC = 1, 2
SORT_TORCH_CUDA, SORT_STABLE, SORT_STABLE_RADIX = 0, 1, 2

EXPORTS = (
    "phnms_abi_version", "phnms_error_string", "phnms_workspace_bytes", "phnms_plan_query", "phnms_plan_query_topk",
    "phnms_forward_f32", "phnms_forward_f32_trace", "phnms_order_workspace_bytes", "phnms_order_f32",
    "phnms_get_lanes_workspace_bytes", "phnms_get_lanes_f32",
    "phnms_decode_lanes_f32", "phnms_line_iou_f32", "phnms_dynamic_k_assign_f32", "phnms_ordered_f64_workspace_bytes", "phnms_forward_ordered_f64", "phnms_forward_collect_f32", "phnms_peer_alloc", "phnms_peer_open", "phnms_peer_close", "phnms_peer_free", "phnms_peer_sync",
)
ABI_VERSION = 5
MAX_DST = 16
IPC_HANDLE_BYTES = 64


class Tuning(ctypes.Structure):
    _fields_ = [("path", ctypes.c_int), ("cluster", ctypes.c_int), ("threads", ctypes.c_int),
                ("max_clusters", ctypes.c_int), ("variant", ctypes.c_int), ("schedule", ctypes.c_int),
                ("stream_warps", ctypes.c_int), ("select_cap", ctypes.c_int), ("lanes_per_pass", ctypes.c_int)]

    def key(self):
        return tuple(getattr(self, k) for k, _ in self._fields_)


class Plan(ctypes.Structure):
    _fields_ = [("path", ctypes.c_int), ("cluster", ctypes.c_int), ("threads", ctypes.c_int),
                ("rows_per_cta", ctypes.c_int), ("smem_bytes", ctypes.c_int), ("grid", ctypes.c_int),
                ("launches", ctypes.c_int), ("variant", ctypes.c_int), ("cols_per_thread", ctypes.c_int), ("max_active_clusters", ctypes.c_int),
                ("workspace_bytes", ctypes.c_size_t)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Collect(ctypes.Structure):
    """phnms_collect: where the compact kept-lane records of a call go (include/phnms.h)."""
    _fields_ = [("n_dst", ctypes.c_int), ("width", ctypes.c_int), ("row0", ctypes.c_int64), ("rows", ctypes.c_int64),
                ("dst", ctypes.c_void_p * MAX_DST),
                ("signal_epoch", ctypes.c_uint64), ("wait_epoch", ctypes.c_uint64), ("timeout_ns", ctypes.c_uint64),
                ("signal_dst", ctypes.c_void_p * MAX_DST), ("wait_src", ctypes.c_void_p), ("status", ctypes.c_void_p),
                ("sync_counter", ctypes.c_void_p)]


def collect(dst_ptrs, rows: int, width: int, row0: int = 0) -> Collect:
    """Destinations are [rows, width] int64 buffers; the call that uses this stores its frames at rows row0 .. row0 + F."""
    if not 1 <= len(dst_ptrs) <= MAX_DST:
        raise ValueError(f"between 1 and {MAX_DST} collection buffers")
    c = Collect()
    c.n_dst, c.width, c.row0, c.rows = len(dst_ptrs), int(width), int(row0), int(rows)
    for i, ptr in enumerate(dst_ptrs):
        c.dst[i] = int(ptr)
    return c


class PhnmsError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"phnms: {msg} (code {code})")
        self.code = code


_lib = None
_shim = None
SHIM_PATH = os.path.join(HERE, "csrc", "nms_impl.so")


def shim():
    """The pybind11 module `nms_impl` (csrc/nms_impl.cpp): the reference's native surface `nms_forward(boxes, scores, thresh,
    top_k)` over libphnms.so, a few microseconds of host time per call.  None when it has not been built (the ctypes path is
    then used; both end in the same C ABI call)."""
    global _shim
    if _shim is None:
        _shim = False
        if os.path.exists(SHIM_PATH) and not os.environ.get("PHNMS_NO_SHIM") and not os.environ.get("PHNMS_SO"):
            import importlib.util
            import torch  # noqa: F401  (libtorch symbols first)
            lib()
            try:
                spec = importlib.util.spec_from_file_location("nms_impl", SHIM_PATH)
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                if mod.abi_version() == ABI_VERSION:
                    _shim = mod
            except Exception:   # noqa: BLE001 -- stale or unloadable build: the ctypes path does the same work
                _shim = False
    return _shim or None


def lib() -> ctypes.CDLL:
    """Load the native library, building it first if the sources are newer and nvcc is present."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        from . import build as _build
        _build.build()
    if not os.path.exists(SO_PATH):
        raise ImportError(f"{SO_PATH} is missing and could not be built; the lane-NMS op has no fallback path")
    L = ctypes.CDLL(SO_PATH)
    vp, i64, ci, sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_size_t
    L.phnms_abi_version.restype = ci
    L.phnms_error_string.argtypes = [ci]
    L.phnms_error_string.restype = ctypes.c_char_p
    L.phnms_workspace_bytes.argtypes = [i64, i64, ci, ctypes.POINTER(Tuning)]
    L.phnms_workspace_bytes.restype = sz
    L.phnms_plan_query.argtypes = [i64, i64, ci, ctypes.POINTER(Tuning), ctypes.POINTER(Plan)]
    L.phnms_plan_query.restype = ci
    L.phnms_plan_query_topk.argtypes = [i64, i64, ci, i64, ctypes.POINTER(Tuning), ctypes.POINTER(Plan)]
    L.phnms_plan_query_topk.restype = ci
    L.phnms_forward_f32.argtypes = [vp, vp, vp, i64, i64, ci, ctypes.c_float, i64, ci, vp, vp, vp, vp, sz,
                                    ctypes.POINTER(Tuning), vp]
    L.phnms_forward_f32.restype = ci
    L.phnms_forward_f32_trace.argtypes = L.phnms_forward_f32.argtypes + [vp, ci]
    L.phnms_forward_f32_trace.restype = ci
    L.phnms_get_lanes_workspace_bytes.argtypes = [i64, i64, ci, ctypes.POINTER(Tuning)]
    L.phnms_get_lanes_workspace_bytes.restype = sz
    L.phnms_get_lanes_f32.argtypes = [vp, i64, i64, ci, ci, ctypes.c_float, ctypes.c_float, ctypes.c_float, i64, ci,
                                      vp, vp, vp, vp, vp, sz, ctypes.POINTER(Tuning), vp]
    L.phnms_get_lanes_f32.restype = ci
    L.phnms_order_workspace_bytes.argtypes = [i64, i64]
    L.phnms_order_workspace_bytes.restype = sz
    L.phnms_order_f32.argtypes = [vp, vp, i64, i64, ci, vp, vp, sz, vp]
    L.phnms_order_f32.restype = ci
    try:
        L.phnms_decode_lanes_f32.argtypes = [vp, vp, i64, i64, ci, ci, vp, ctypes.c_double, ctypes.c_double, vp, vp, vp, vp]
        L.phnms_decode_lanes_f32.restype = ci
        L.phnms_line_iou_f32.argtypes = [vp, vp, i64, i64, ci, ctypes.c_float, ctypes.c_float, ci, vp, vp]
        L.phnms_line_iou_f32.restype = ci
        L.phnms_dynamic_k_assign_f32.argtypes = [vp, vp, i64, i64, i64, ci, ci, ci, ctypes.c_float, vp, vp, vp, vp]
        L.phnms_dynamic_k_assign_f32.restype = ci
        L.phnms_ordered_f64_workspace_bytes.argtypes = [i64, i64]
        L.phnms_ordered_f64_workspace_bytes.restype = sz
        L.phnms_forward_ordered_f64.argtypes = [vp, vp, vp, i64, i64, ci, ctypes.c_float, i64, vp, vp, vp, vp, sz, vp]
        L.phnms_forward_ordered_f64.restype = ci
        L.phnms_forward_collect_f32.argtypes = L.phnms_forward_f32.argtypes + [ctypes.POINTER(Collect)]
        L.phnms_forward_collect_f32.restype = ci
        L.phnms_peer_alloc.argtypes = [sz, ctypes.POINTER(vp), ctypes.c_char_p]
        L.phnms_peer_alloc.restype = ci
        L.phnms_peer_open.argtypes = [ctypes.c_char_p, ctypes.POINTER(vp)]
        L.phnms_peer_open.restype = ci
        L.phnms_peer_close.argtypes = [vp]
        L.phnms_peer_close.restype = ci
        L.phnms_peer_free.argtypes = [vp]
        L.phnms_peer_free.restype = ci
        L.phnms_peer_sync.argtypes = [ctypes.POINTER(vp), vp, ci, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, vp, vp]
        L.phnms_peer_sync.restype = ci
    except AttributeError:      # an older build loaded through PHNMS_SO for A/B timing: no collection entry points
        if not os.environ.get("PHNMS_SO"):
            raise
    if L.phnms_abi_version() != ABI_VERSION and not os.environ.get("PHNMS_SO"):
        raise ImportError("libphnms.so ABI version mismatch; rebuild with `python -m phnet_b200.build`")
    _lib = L
    return L


def check(code: int) -> None:
    if code != 0:
        raise PhnmsError(code, lib().phnms_error_string(code).decode())


def tuning(path: int = 0, cluster: int = 0, threads: int = 0, max_clusters: int = 0, variant: int = 0, schedule: int = 0,
           stream_warps: int = 0, select_cap: int = 0, lanes_per_pass: int = 0):
    if not (path or cluster or threads or max_clusters or variant or schedule or stream_warps or select_cap or lanes_per_pass):
        return None
    return Tuning(path, cluster, threads, max_clusters, variant, schedule, stream_warps, select_cap, lanes_per_pass)


def plan(F: int, N: int, n_off: int, tune: Tuning | None = None, top_k: int = -1) -> dict:
    p = Plan()
    check(lib().phnms_plan_query_topk(F, N, n_off, int(top_k), ctypes.byref(tune) if tune else None, ctypes.byref(p)))
    return p.as_dict()

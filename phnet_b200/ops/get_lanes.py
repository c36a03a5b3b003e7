"""Clip-level `get_lanes` post-processing on the device.

Mirrors the tensor part of PHNet's `get_lanes` (libs/models/Router4OL.py:437-479, Router4OLV2.py:406-448,
RouterV4.py:394-442) for a whole clip at once: confidence filter, column drop + pixel/strip scaling, lane NMS,
`predictions[keep]`, rounding of the length column(s).  The reference runs this once per frame in Python with two host
syncs; here T frames take four launches (prepare, top-M select, fused NMS, gather) and no sync.  `decode_lanes` then does
the tensor part of `predictions_to_pred` (the `points` of every `Lane`) in one more launch; only the scipy spline inside
`Lane` (libs/utils/lane.py:4-16) stays on the host.
"""
from __future__ import annotations

import ctypes

import torch

from .. import _capi

__all__ = ["get_lanes", "decode_lanes"]


def get_lanes(output: torch.Tensor, conf_threshold: float, nms_thres: float, max_lanes: int, img_w: int = 768, *,
              sort_model: int = _capi.SORT_TORCH_CUDA, tuning=None):
    """output [T, A, 6 + n_off] (OpenLane-V heads) or [T, A, 7 + n_off] (VIL-100 heads, pass them with `vil=True` via the
    last dimension being 7 + 36): raw per-prior predictions of T frames, fp32 CUDA contiguous.

    Returns (lanes[T, max_lanes, C], num[T] int64, index[T, max_lanes] int64, keep_inds[T, A] bool):
    frame t's kept predictions are lanes[t, :num[t]] -- exactly `predictions` after line Router4OL.py:470 -- and
    index[t, :num[t]] are their prior indices in the unfiltered frame."""
    if not output.is_cuda or output.dtype != torch.float32 or not output.is_contiguous() or output.dim() != 3:
        raise RuntimeError("output must be a contiguous float32 CUDA tensor [T, A, hdr + n_off]")
    T, A, C = output.shape
    if C - 6 in (36, 72):
        hdr = 6
    elif C - 7 in (36, 72):
        hdr = 7
    else:
        raise RuntimeError("rows must be 6 + n_off (OpenLane-V) or 7 + n_off (VIL-100) wide with n_off in {36, 72}")
    n_off = C - hdr
    K = int(max_lanes)
    dev = output.device
    lanes = torch.empty((T, K, C), dtype=torch.float32, device=dev)
    num = torch.empty((T,), dtype=torch.int64, device=dev)
    index = torch.empty((T, K), dtype=torch.int64, device=dev)
    keep_inds = torch.empty((T, A), dtype=torch.uint8, device=dev)
    L = _capi.lib()
    t = tuning if isinstance(tuning, _capi.Tuning) or tuning is None else _capi.tuning(**tuning)
    tp = ctypes.byref(t) if t is not None else None
    nbytes = L.phnms_get_lanes_workspace_bytes(T, A, n_off, tp)
    ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = L.phnms_get_lanes_f32(output.data_ptr(), T, A, n_off, hdr, float(conf_threshold), float(img_w),
                                   float(nms_thres), K, int(sort_model), lanes.data_ptr(), num.data_ptr(),
                                   index.data_ptr(), keep_inds.data_ptr(), ws.data_ptr(), nbytes, tp,
                                   torch.cuda.current_stream().cuda_stream)
    _capi.check(rc)
    ws.record_stream(torch.cuda.current_stream(dev))
    return lanes, num, index, keep_inds.bool()


def decode_lanes(lanes: torch.Tensor, num: torch.Tensor, ori_img_h: float = 1.0, cut_height: float = 0.0,
                 prior_ys: torch.Tensor | None = None):
    """`predictions_to_pred` for a whole clip on the device (libs/models/Router4OLV2.py:363-404 for 6 + n_off wide rows,
    RouterV4.py:349-392 for 7 + n_off wide rows): the `points` arrays the reference hands to `Lane(points=...)`.

    lanes [T, K, C], num [T]: what `get_lanes` returned.  prior_ys: the model's buffer (default: torch.linspace(1, 0, n_off)).
    Returns (points[T, K, n_off, 2] float64, npoints[T, K] int32, meta[T, K, 3] float32 = start_x, start_y, conf);
    lane k of frame t has points[t, k, :npoints[t, k]]; npoints == 0 where the reference drops the lane (<= 1 point)."""
    if not lanes.is_cuda or lanes.dtype != torch.float32 or not lanes.is_contiguous() or lanes.dim() != 3:
        raise RuntimeError("lanes must be a contiguous float32 CUDA tensor [T, K, hdr + n_off]")
    T, K, C = lanes.shape
    if C - 6 in (36, 72):
        hdr = 6
    elif C - 7 in (36, 72):
        hdr = 7
    else:
        raise RuntimeError("rows must be 6 + n_off (OpenLane-V) or 7 + n_off (VIL-100) wide with n_off in {36, 72}")
    n_off = C - hdr
    dev = lanes.device
    if num.device != dev or num.dtype != torch.int64 or num.shape != (T,):
        raise RuntimeError("num must be an int64 tensor [T] on the same device")
    if prior_ys is None:
        prior_ys = torch.linspace(1, 0, steps=n_off, dtype=torch.float32)        # Router4OLV2.py:61
    ys = prior_ys.to(device=dev, dtype=torch.float64).contiguous()               # `.double()`, Router4OLV2.py:368
    if ys.numel() != n_off:
        raise RuntimeError("prior_ys must have n_off entries")
    points = torch.empty((T, K, n_off, 2), dtype=torch.float64, device=dev)
    npoints = torch.empty((T, K), dtype=torch.int32, device=dev)
    meta = torch.empty((T, K, 3), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _capi.lib().phnms_decode_lanes_f32(lanes.data_ptr(), num.data_ptr(), T, K, n_off, hdr, ys.data_ptr(),
                                                float(ori_img_h), float(cut_height), points.data_ptr(), npoints.data_ptr(),
                                                meta.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _capi.check(rc)
    ys.record_stream(torch.cuda.current_stream(dev))
    return points, npoints, meta

"""Host-side mirror of PHNet's `libs/ops` lane NMS, backed by libphnms.so (hand-written sm_100a CUDA).

Reference interface mirrored (paths relative to the PHNet repo):
    libs/ops/nms.py:32-33          def nms(boxes, scores, overlap, top_k)
    libs/ops/csrc/nms.cpp:44-57    nms_forward: sort scores descending, CUDA + contiguity checks
    libs/ops/csrc/nms_kernel.cu:147-192  nms_cuda_forward: shape checks, outputs keep / num_to_keep / parent_object_index
Callers: the five `get_lanes` copies (libs/models/Router4OL.py:460-465 and siblings) do

    keep, num_to_keep, _ = nms(nms_predictions, scores, overlap=nms_thres, top_k=nms_topk)
    keep = keep[:num_to_keep]

so the keyword names `overlap` / `top_k`, the list-of-three return value, int64 dtype, CUDA device, the 0-dim
`num_to_keep` and the zero padding of `keep` are all part of the contract kept here.

PyTorch is used for device memory and the current stream only; there is no fallback: if the native library
cannot be loaded, or the tensors are not on a CUDA device, the op raises.
"""
from __future__ import annotations

import ctypes

import torch

from .. import _capi

__all__ = ["nms", "nms_batched", "sort_order", "plan"]


def _check_inputs(boxes: torch.Tensor, scores: torch.Tensor, batched: bool):
    # error behaviour of the reference: CHECK_CUDA / CHECK_CONTIGUOUS (nms.cpp:40-42,53-54; nms_kernel.cu:167),
    # AT_DISPATCH_FLOATING_TYPES (nms_kernel.cu:171), row-width check (:154)
    if not isinstance(boxes, torch.Tensor) or not isinstance(scores, torch.Tensor):
        raise TypeError("nms: boxes and scores must be torch tensors")
    if not boxes.is_cuda:
        raise RuntimeError("boxes must be a CUDA tensor")
    if not scores.is_cuda:
        raise RuntimeError("scores must be a CUDA tensor")
    if scores.device != boxes.device:
        raise RuntimeError("boxes and scores must be on the same device")
    if not boxes.is_contiguous():
        raise RuntimeError("boxes must be contiguous")
    if boxes.dtype not in (torch.float32, torch.float64):     # AT_DISPATCH_FLOATING_TYPES: float and double only
        raise RuntimeError(f"\"nms_cuda_forward\" not implemented for '{boxes.dtype}'")
    want = 3 if batched else 2
    if boxes.dim() != want:
        raise RuntimeError(f"boxes must have {want} dimensions, got {boxes.dim()}")
    if boxes.shape[-1] < 6:
        raise RuntimeError("Wrong number of offsets. Rows are 5 + n_offsets wide")
    if scores.dtype != torch.float32 and boxes.dtype == torch.float32:
        # The reference sorts whatever dtype it is given (`scores.sort(0, True)`, nms.cpp:51); a cast to fp32 could merge scores
        # that differ only beyond fp32 precision and change the order.  So: torch's own sort of the original dtype, then fp32
        # stand-in scores that encode that order exactly (-rank: distinct, exact below 2^24 entries).
        order = scores.sort(-1, True)[1]
        ranks = torch.empty_like(order)
        ranks.scatter_(-1, order, torch.arange(scores.shape[-1], device=scores.device).expand_as(order).contiguous())
        scores = -ranks.to(torch.float32)
    if not scores.is_contiguous():
        scores = scores.contiguous()     # `scores.sort` accepts strided input (nms.cpp:51)
    if scores.shape != boxes.shape[:-1]:
        raise RuntimeError("scores must have one entry per proposal")
    return boxes, scores


def _as_top_k(top_k) -> int:
    top_k = int(top_k)
    if top_k < 0:
        raise TypeError("top_k must be non-negative (unsigned long in the reference, nms.cpp:48)")
    return top_k


_ws_bytes_cache: dict = {}     # (F, N, n_off, tuning key) -> workspace bytes (a pure function of the shape)
_ws_cache: dict = {}           # (device index, stream handle) -> reusable workspace tensor for small calls


def _workspace(L, dev, stream, F, N, n_off, t, tp):
    key = (F, N, n_off, None if t is None else t.key())
    nbytes = _ws_bytes_cache.get(key)
    if nbytes is None:
        nbytes = _ws_bytes_cache[key] = int(L.phnms_workspace_bytes(F, N, n_off, tp))
    if nbytes == 0:
        return None, 0
    if nbytes > (1 << 20):                      # large batches: a fresh allocation (stream-ordered by the caching allocator)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        ws.record_stream(torch.cuda.current_stream(dev))
        return ws, nbytes
    # per-frame calls: one small workspace per (device, stream), reused -- launches on one stream are ordered
    ck = (dev.index, stream)
    ws = _ws_cache.get(ck)
    if ws is None or ws.numel() < nbytes:
        ws = _ws_cache[ck] = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8, device=dev)
    return ws, nbytes


def _launch(boxes, scores, n_valid, F, N, n_off, overlap, top_k, sort_model, tune, keep, num, parent, collect=None):
    top_k = int(top_k)
    if top_k < 0:
        raise TypeError("top_k must be non-negative (unsigned long in the reference, nms.cpp:48)")
    L = _capi.lib()
    t = tune if isinstance(tune, _capi.Tuning) or tune is None else _capi.tuning(**tune)
    tp = ctypes.byref(t) if t is not None else None
    dev = boxes.device
    switch = torch.cuda.current_device() != dev.index
    if switch:
        prev = torch.cuda.current_device()
        torch.cuda.set_device(dev)
    try:
        stream = torch.cuda.current_stream(dev).cuda_stream
        ws, ws_bytes = _workspace(L, dev, stream, F, N, n_off, t, tp)
        args = (boxes.data_ptr(), scores.data_ptr(), n_valid.data_ptr() if n_valid is not None else None,
                F, N, n_off, float(overlap), top_k, int(sort_model), keep.data_ptr(), num.data_ptr(),
                parent.data_ptr(), ws.data_ptr() if ws is not None else None, ws_bytes, tp, stream)
        if collect is None:
            rc = L.phnms_forward_f32(*args)
        else:
            rc = L.phnms_forward_collect_f32(*args, ctypes.byref(collect))
    finally:
        if switch:
            torch.cuda.set_device(prev)
    _capi.check(rc)


def _launch_f64(boxes, scores, n_valid, F, N, n_off, overlap, top_k, keep, num, parent):
    """Double boxes (the reference dispatches over float and double, nms_kernel.cu:171): the ordering is torch's own
    `scores.sort(..., descending=True)` -- literally what the reference calls (nms.cpp:51), so ties fall the same way --
    and the bitmask + scan kernels of the C ABI do the rest in fp64."""
    top_k = int(top_k)
    if top_k < 0:
        raise TypeError("top_k must be non-negative (unsigned long in the reference, nms.cpp:48)")
    L = _capi.lib()
    dev = boxes.device
    with torch.cuda.device(dev):
        if n_valid is not None:
            # ragged batch: the padding sorts behind every real proposal (-inf; among equal scores torch's sort of more than 32
            # elements is stable, so a real -inf score still precedes the padding), frame f is then ordered like
            # `scores[f, :n_valid[f]].sort(0, True)` (frames of <= 32 real proposals: torch's unstable small sort may break ties
            # differently on the padded row than on the slice)
            pad = torch.arange(N, device=dev)[None, :] >= n_valid.to(torch.int64)[:, None]
            scores = scores.reshape(F, N).masked_fill(pad, float("-inf"))
        order = scores.sort(-1, True)[1].contiguous()
        nbytes = int(L.phnms_ordered_f64_workspace_bytes(F, N))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        rc = L.phnms_forward_ordered_f64(boxes.data_ptr(), order.data_ptr(), n_valid.data_ptr() if n_valid is not None else None,
                                         F, N, n_off, float(overlap), top_k,
                                         keep.data_ptr(), num.data_ptr(), parent.data_ptr(), ws.data_ptr(), nbytes,
                                         torch.cuda.current_stream(dev).cuda_stream)
        ws.record_stream(torch.cuda.current_stream(dev))
        order.record_stream(torch.cuda.current_stream(dev))
    _capi.check(rc)


_shim_forward = None     # nms_impl.nms_forward once bound; False when the shim is not built


def _bind_shim():
    global _shim_forward
    sh = _capi.shim()
    _shim_forward = getattr(sh, "nms_forward_or_none", False) if sh is not None else False
    return _shim_forward


def nms(boxes: torch.Tensor, scores: torch.Tensor, overlap, top_k, *, sort_model: int = _capi.SORT_TORCH_CUDA,
        tuning=None):
    """Drop-in for `libs.ops.nms` (libs/ops/nms.py:32).

    boxes  [N, 5+n_off] fp32 CUDA contiguous; scores [N]; overlap: pixel threshold; top_k: stop after this many lanes.
    Returns [keep[N] int64, num_to_keep[] int64, parent_object_index[N] int64] on the same device.
    """
    # the per-frame call PHNet makes: straight to the pybind shim -- the reference's own native signature (nms.cpp:44-48), same
    # C ABI call.  The shim itself decides whether the arguments are the plain float32 case (None otherwise): every attribute
    # lookup in Python is ~0.1 us of a ~7 us call.
    if tuning is None and sort_model == 0:
        fwd = _shim_forward if _shim_forward is not None else _bind_shim()
        if fwd is not False:
            try:
                out = fwd(boxes, scores, overlap, top_k)
                if out is not None:
                    return out
            except TypeError:
                pass    # argument types the pybind signature does not take (numpy scalars, tensor subclasses ...): the general path
    boxes, scores = _check_inputs(boxes, scores, batched=False)
    N, P = boxes.shape
    out = torch.empty(2 * N + 1, dtype=torch.int64, device=boxes.device)   # one allocation, three views
    keep, parent, num = out[:N], out[N:2 * N], out[2 * N]
    if boxes.dtype == torch.float64:
        _launch_f64(boxes, scores, None, 1, N, P - 5, overlap, top_k, keep, num, parent)
    else:
        _launch(boxes, scores, None, 1, N, P - 5, overlap, top_k, sort_model, tuning, keep, num, parent)
    return [keep, num, parent]


def nms_batched(boxes: torch.Tensor, scores: torch.Tensor, overlap, top_k, n_valid: torch.Tensor | None = None, *,
                sort_model: int = _capi.SORT_TORCH_CUDA, tuning=None, out=None, collect=None):
    """F independent `nms` calls in one launch.

    boxes [F, N, 5+n_off], scores [F, N], n_valid [F] int32 (optional: real proposals per frame, rest is padding).
    Returns (keep[F, N], num_to_keep[F], parent_object_index[F, N]); frame f equals
    `nms(boxes[f, :n_valid[f]], scores[f, :n_valid[f]], overlap, top_k)` padded with zeros to N.

    collect: optional `_capi.Collect` (see `phnet_b200.peer`): the kernels additionally store the compact record
    {keep[f, :top_k], num_to_keep[f]} of every frame into each listed buffer -- local tensors or other GPUs' memory.
    """
    boxes, scores = _check_inputs(boxes, scores, batched=True)
    F, N, P = boxes.shape
    dev = boxes.device
    if n_valid is not None:
        if n_valid.device != dev or n_valid.dtype != torch.int32 or n_valid.shape != (F,) or not n_valid.is_contiguous():
            raise RuntimeError("n_valid must be a contiguous int32 tensor of shape [F] on the same device")
    if out is None:
        keep = torch.empty((F, N), dtype=torch.int64, device=dev)
        num = torch.empty((F,), dtype=torch.int64, device=dev)
        parent = torch.empty((F, N), dtype=torch.int64, device=dev)
    else:
        keep, num, parent = out
    if boxes.dtype == torch.float64:
        if collect is not None:
            raise RuntimeError("collect is not supported for float64 boxes")
        _launch_f64(boxes, scores, n_valid, F, N, P - 5, overlap, top_k, keep, num, parent)
    else:
        _launch(boxes, scores, n_valid, F, N, P - 5, overlap, top_k, sort_model, tuning, keep, num, parent, collect)
    return keep, num, parent


def sort_order(scores: torch.Tensor, n_valid: torch.Tensor | None = None, *, sort_model: int = _capi.SORT_TORCH_CUDA):
    """The ordering step alone: `scores.sort(0, True)[1]` of libs/ops/csrc/nms.cpp:51, per frame.  scores [F, N] or [N]."""
    if not scores.is_cuda:
        raise RuntimeError("scores must be a CUDA tensor")
    one = scores.dim() == 1
    s = scores.reshape(1, -1) if one else scores
    s = s.contiguous().float()
    F, N = s.shape
    L = _capi.lib()
    order = torch.zeros((F, N), dtype=torch.int64, device=s.device)
    ws_bytes = L.phnms_order_workspace_bytes(F, N)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=s.device)
    with torch.cuda.device(s.device):
        rc = L.phnms_order_f32(s.data_ptr(), n_valid.data_ptr() if n_valid is not None else None, F, N, int(sort_model),
                               order.data_ptr(), ws.data_ptr(), ws_bytes, torch.cuda.current_stream().cuda_stream)
    _capi.check(rc)
    ws.record_stream(torch.cuda.current_stream(s.device))
    return order[0] if one else order


def plan(F: int, N: int, n_off: int, tuning=None, top_k: int = -1) -> dict:
    """What one call would launch for this shape (path, variant, threads, shared memory, grid); top_k < 0: PHNet's 1..8."""
    t = tuning if isinstance(tuning, _capi.Tuning) or tuning is None else _capi.tuning(**tuning)
    return _capi.plan(F, N, n_off, t, top_k)

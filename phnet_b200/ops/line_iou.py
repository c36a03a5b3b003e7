"""`line_iou` of PHNet's training code (libs/utils/dynamic_assign.py:5-36) on the device: SURVEY.md section 8f row 4.

Same signature and meaning as the reference function: `line_iou(pred, target, img_w, length=15, aligned=True)`;
aligned=True returns the per-pair IoU [n] (the LIoU loss is `1 - line_iou(...)`, :38-42), aligned=False the pairwise
[num_pred, num_target] matrix used by the dynamic-k assignment (:83-125).  Forward only (no autograd): training stays out
of scope, this is the adjacent component named "next" in the scope table.
"""
from __future__ import annotations

import torch

from .. import _capi

__all__ = ["line_iou"]


def line_iou(pred: torch.Tensor, target: torch.Tensor, img_w, length=15, aligned=True) -> torch.Tensor:
    if not (pred.is_cuda and target.is_cuda) or pred.device != target.device:
        raise RuntimeError("pred and target must be CUDA tensors on the same device")
    if pred.dim() != 2 or target.dim() != 2 or pred.shape[1] != target.shape[1]:
        raise RuntimeError("pred [num_pred, n_off] and target [num_target, n_off] must have the same number of offsets")
    if aligned and pred.shape[0] != target.shape[0]:
        raise RuntimeError("aligned=True needs as many predictions as targets")
    p = pred.detach().to(torch.float32).contiguous()
    t = target.detach().to(torch.float32).contiguous()
    out = torch.empty((p.shape[0],) if aligned else (p.shape[0], t.shape[0]), dtype=torch.float32, device=p.device)
    with torch.cuda.device(p.device):
        rc = _capi.lib().phnms_line_iou_f32(p.data_ptr(), t.data_ptr(), p.shape[0], t.shape[0], p.shape[1], float(img_w),
                                            float(length), 1 if aligned else 0, out.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream)
    _capi.check(rc)
    return out

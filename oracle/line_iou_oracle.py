"""CPU restatement of PHNet's training-side `line_iou` (TEST INFRASTRUCTURE: imported by tests/ only).

Follows libs/utils/dynamic_assign.py:5-36 in float32 numpy (sequential float32 accumulation over the offsets).  Pinned
against the reference function itself, imported from /root/reference by tests/golden/make_line_iou_fixtures.py
(fixtures tests/golden/line_iou_ref.npz; tolerance 1e-5 relative: torch reduces over the offsets in a different order).
"""
from __future__ import annotations

import numpy as np


def line_iou(pred, target, img_w, length=15, aligned=True):
    pred = np.asarray(pred, dtype=np.float32)
    target = np.asarray(target, dtype=np.float32)
    length = np.float32(length)
    px1, px2 = pred - length, pred + length                       # :15-18
    tx1, tx2 = target - length, target + length
    if aligned:                                                   # :20-23
        invalid = target
        ovr = np.minimum(px2, tx2) - np.maximum(px1, tx1)
        union = np.maximum(px2, tx2) - np.minimum(px1, tx1)
    else:                                                         # :24-30
        invalid = np.broadcast_to(target[None], (pred.shape[0],) + target.shape)
        ovr = np.minimum(px2[:, None, :], tx2[None]) - np.maximum(px1[:, None, :], tx1[None])
        union = np.maximum(px2[:, None, :], tx2[None]) - np.minimum(px1[:, None, :], tx1[None])
    bad = (invalid < 0) | (invalid >= img_w)                      # :32-34
    ovr = np.where(bad, np.float32(0), ovr).astype(np.float32)
    union = np.where(bad, np.float32(0), union).astype(np.float32)
    so = np.zeros(ovr.shape[:-1], dtype=np.float32)
    su = np.zeros(ovr.shape[:-1], dtype=np.float32)
    for i in range(ovr.shape[-1]):
        so = so + ovr[..., i]
        su = su + union[..., i]
    return so / (su + np.float32(1e-9))                           # :35

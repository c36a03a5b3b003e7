"""CPU restatement of PHNet's `predictions_to_pred` (TEST INFRASTRUCTURE: imported by tests/ only).

Follows, statement by statement,
    libs/models/Router4OLV2.py:363-404   (OpenLane-V models, rows = 2 scores, start_y, start_x, theta, length, x[n_off])
    libs/models/RouterV4.py:349-392      (VIL-100 models, rows additionally carry invalid_len before the x[n_off])
and returns, per kept lane, the `points` array the reference passes to `Lane(points=...)` (libs/utils/lane.py:4-16) plus
the metadata triple.  Pinned against the reference's own function bodies, executed from /root/reference by
tests/golden/make_decode_fixtures.py (fixtures in tests/golden/decode_ref.npz, checked in tests/test_oracle.py).
"""
from __future__ import annotations

import numpy as np


def prior_ys(n_off: int) -> np.ndarray:
    """`torch.linspace(1, 0, steps=n_off, dtype=float32).double()` (Router4OLV2.py:61,368)."""
    import torch
    return torch.linspace(1, 0, steps=n_off, dtype=torch.float32).double().numpy()


def predictions_to_pred(predictions: np.ndarray, hdr: int, ori_img_h: float = 1.0, cut_height: float = 0.0, ys=None):
    """predictions [L, hdr + n_off] float32 (after the length rounding of get_lanes).  Returns a list with one entry per
    lane: (points [n, 2] float64, (start_x, start_y, conf)) or None where the reference skips the lane."""
    pred = np.array(predictions, dtype=np.float32, copy=True)
    n_off = pred.shape[1] - hdr
    n_strips = n_off - 1
    ys = prior_ys(n_off) if ys is None else np.asarray(ys, dtype=np.float64)
    out = []
    for lane in pred:
        lane_xs = lane[hdr:]                                                   # a view, modified in place like the reference
        start = min(max(0, int(round(float(lane[2]) * n_strips))), n_strips)   # Router4OLV2.py:373-374
        if hdr == 7:
            start += int(round(float(lane[6])))                                # RouterV4.py:360-362
        length = int(round(float(lane[5])))
        end = start + length - 1
        end = min(end, len(ys) - 1)
        if hdr == 6:                                                           # Router4OLV2.py:382-385
            head = lane_xs[:start]
            mask = ~(((head >= 0.) & (head <= 1.))[::-1].cumprod()[::-1].astype(bool))
            lane_xs[end + 1:] = -2
            lane_xs[:start][mask] = -2
        else:                                                                  # RouterV4.py:370-372
            lane_xs[end + 1:] = -2
            lane_xs[:start] = -2
        sel = lane_xs >= 0
        lys = ys[sel]
        lxs = lane_xs[sel][::-1].astype(np.float64)
        lys = lys[::-1]
        if hdr == 7:
            lys = (lys * (ori_img_h - cut_height) + cut_height) / ori_img_h    # RouterV4.py:378
        if len(lxs) <= 1:
            out.append(None)
            continue
        out.append((np.stack((lxs, lys), axis=1), (float(lane[3]), float(lane[2]), float(lane[1]))))
    return out

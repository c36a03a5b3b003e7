"""On-disk result format of PHNet's test scripts, written from the batched device decode.

Mirrors `evaluation/generate_lane.py:31-61`: one `<ImgName>.lines.txt` per frame under `<root>/<clip name>/`, one text line
per lane with more than two points, the points in reversed `Lane.points` order:
    generate_pred   (VIL-100,    testVIL.py):   '%d %d '     % (tx * W, ty * H)
    generate_predV2 (OpenLane-V, testOLV3.py):  '%.1f %.1f ' % (tx * W / 2, (ty * H + 480) / 2)
with (H, W) = info['size'].  These files are what `evaluation/culane` (the C++ evaluator) reads, so accuracy can be
re-checked end to end with the reference's own tools.  Input: the arrays `phnet_b200.ops.decode_lanes` returns, moved to
the host once per clip (one D2H copy instead of the reference's several `.cpu()` / `.item()` syncs per lane).
"""
from __future__ import annotations

import os

import numpy as np

FORMATS = ("vil", "openlane")


def frame_lines(points: np.ndarray, npoints: np.ndarray, size, fmt: str) -> str:
    """Text of one frame's file.  points [K, n_off, 2] float64, npoints [K]; size = (H, W) as in info['size']."""
    if fmt not in FORMATS:
        raise ValueError(f"fmt must be one of {FORMATS}")
    h, w = size[0], size[1]
    out = []
    for k in range(points.shape[0]):
        n = int(npoints[k])
        if n > 2:                                            # `if len(lane.points) > 2`, generate_lane.py:41,56
            for tx, ty in points[k, :n][::-1]:               # `reversed(lane.points)`
                if fmt == "vil":
                    out.append('%d %d ' % (tx * w, ty * h))                            # :43
                else:
                    out.append('%.1f %.1f ' % (tx * w / 2, (ty * h + 480) / 2))        # :60
            out.append('\n')
    return "".join(out)


def write_clip(points, npoints, root: str, clip_name: str, img_names, size, fmt: str):
    """points [T, K, n_off, 2], npoints [T, K] (torch tensors on any device, or numpy): writes T files and returns their
    paths.  Directory layout and file names as generate_pred / generate_predV2 produce them."""
    if hasattr(points, "detach"):
        points = points.detach().cpu().numpy()
    if hasattr(npoints, "detach"):
        npoints = npoints.detach().cpu().numpy()
    if len(img_names) != points.shape[0]:
        raise ValueError("one image name per frame")
    d = os.path.join(root, clip_name)
    os.makedirs(d, exist_ok=True)
    paths = []
    for t, name in enumerate(img_names):
        path = os.path.join(d, name + '.lines.txt')
        with open(path, "w") as fp:
            fp.write(frame_lines(points[t], npoints[t], size, fmt))
        paths.append(path)
    return paths

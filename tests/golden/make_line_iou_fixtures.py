"""Generates tests/golden/line_iou_ref.npz by IMPORTING the reference's `line_iou` (libs/utils/dynamic_assign.py:5-36) from
/root/reference and running it with CPU torch on seeded inputs.

    python tests/golden/make_line_iou_fixtures.py        (needs /root/reference; CPU only)
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from libs.utils.dynamic_assign import line_iou  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "line_iou_ref.npz")


def lanes(n, n_off, g, img_w):
    base = torch.rand(n, 1, generator=g) * img_w
    slope = (torch.rand(n, 1, generator=g) - 0.5) * 8
    x = base + slope * torch.arange(n_off, dtype=torch.float32)[None] + torch.randn(n, n_off, generator=g) * 5
    return x


def main():
    out = {}
    for case, (n_off, img_w, length, npred, ntgt) in enumerate(((72, 768, 15, 240, 5), (36, 768, 15, 240, 8), (72, 1920, 30, 300, 40),
                                                                (72, 768, 15, 1, 1), (36, 640, 7.5, 130, 33))):
        g = torch.Generator().manual_seed(100 + case)
        pred, tgt = lanes(npred, n_off, g, img_w), lanes(ntgt, n_off, g, img_w)
        tgt[:, :3] = -1e5                                  # the reference marks offsets without ground truth as far negative
        tgt[0, -4:] = float(img_w)                         # exactly img_w is invalid, img_w - 1 is valid
        if ntgt > 1:
            tgt[1, 5] = img_w - 1.0
            tgt[1] = pred[0]                               # identical lanes: IoU 1 where valid
        out[f"c{case}_pred"], out[f"c{case}_tgt"] = pred.numpy(), tgt.numpy()
        out[f"c{case}_par"] = np.array([img_w, length], dtype=np.float64)
        out[f"c{case}_pair"] = line_iou(pred.clone(), tgt.clone(), img_w, length=length, aligned=False).numpy()
        n = min(npred, ntgt)
        out[f"c{case}_aligned"] = line_iou(pred[:n].clone(), tgt[:n].clone(), img_w, length=length, aligned=True).numpy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT)


if __name__ == "__main__":
    main()

"""Generates tests/golden/dynamic_assign_ref.npz by IMPORTING the reference's `dynamic_k_assign` and `dynamic_k_assign_CF`
(libs/utils/dynamic_assign.py:83-125, :327-370) and `line_iou` (:5-36) from /root/reference and running them with CPU torch on
seeded inputs: cost / IoU matrices built the way `anc_assign` builds them (:212-247: a score product plus a focal term, minus
nothing; IoUs from line_iou on clustered lanes), plain random matrices, more ground truths than the priors can serve, and a case in
which priors already taken are matched again (their INFINITY beats every remaining cost) so that the "matched to several ground
truths" branch (:116-120) runs.

    python tests/golden/make_dynamic_assign_fixtures.py        (needs /root/reference; CPU only)
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from libs.utils.dynamic_assign import dynamic_k_assign, dynamic_k_assign_CF, line_iou  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dynamic_assign_ref.npz")


def lanes(n, n_off, g, img_w, base=None):
    b = torch.rand(n, 1, generator=g) * img_w if base is None else base
    slope = (torch.rand(n, 1, generator=g) - 0.5) * 6
    return b + slope * torch.arange(n_off, dtype=torch.float32)[None] + torch.randn(n, n_off, generator=g) * 4


def main():
    out = {}
    case = 0

    def emit(cost, iou, cf=False):
        nonlocal case
        fn = dynamic_k_assign_CF if cf else dynamic_k_assign
        p, g = fn(cost.clone(), iou.clone())
        out[f"c{case}_cost"], out[f"c{case}_iou"] = cost.numpy(), iou.numpy()
        out[f"c{case}_cf"] = np.array([1 if cf else 0])
        out[f"c{case}_prior"], out[f"c{case}_gt"] = p.numpy().astype(np.int64), g.numpy().astype(np.int64)
        case += 1

    # (a) lane-shaped problems: priors clustered around the ground truths, IoU from the reference's own line_iou
    for n_off, img_w, npri, ngt in ((72, 768, 240, 4), (36, 768, 240, 8), (72, 768, 240, 1), (72, 768, 100, 3), (36, 640, 33, 5),
                                    (72, 768, 1000, 12)):
        g = torch.Generator().manual_seed(500 + case)
        tgt = lanes(ngt, n_off, g, img_w)
        which = torch.randint(0, ngt, (npri,), generator=g)
        pred = lanes(npri, n_off, g, img_w, base=tgt[which][:, :1] + torch.randn(npri, 1, generator=g) * 25)
        iou = line_iou(pred, tgt, img_w, length=12, aligned=False)
        dist = (pred[:, None, :] - tgt[None]).abs().mean(-1)
        score = 1 - dist / dist.max() + 1e-2
        cost = -(score ** 2) * 3.0 + torch.rand(npri, ngt, generator=g) * 0.5          # like anc_assign's cost (:243-244)
        emit(cost, iou)
        emit(cost, iou, cf=True)
    # (b) plain random matrices, IoUs partly negative
    for npri, ngt in ((240, 6), (64, 2), (4, 1), (60, 9)):   # (never more picks than priors: ties among taken rows are torch.topk's to break)
        g = torch.Generator().manual_seed(600 + case)
        emit(torch.randn(npri, ngt, generator=g), torch.rand(npri, ngt, generator=g) * 1.6 - 0.4)
        emit(torch.randn(npri, ngt, generator=g), torch.rand(npri, ngt, generator=g) * 1.2, cf=True)
    # (c) rows taken by an earlier column are picked again: every other cost exceeds INFINITY (987654), the taken rows are the
    #     complete set of smallest entries, so torch.topk has no choice to make; the priors end up with two ground truths
    g = torch.Generator().manual_seed(700)
    npri, ngt = 12, 3
    cost = torch.rand(npri, ngt, generator=g) * 10 + 1.0e6
    cost[3, 0], cost[7, 0] = -5.0, -4.0                      # column 0 takes priors 3 and 7 (k = 2)
    cost[3, 1], cost[7, 1] = 1.0e6 + 50.0, 1.0e6 + 40.0      # ... original costs decide which ground truth they keep
    cost[3, 2], cost[7, 2] = -7.0, 1.0e6 + 60.0
    iou = torch.zeros(npri, ngt)
    iou[:4, :] = 0.55                                        # top-4 sum 2.2 -> k = 2 for every column
    emit(cost, iou)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, case, "cases")


if __name__ == "__main__":
    main()

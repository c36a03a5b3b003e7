"""CPU restatement of PHNet's training-side `dynamic_k_assign` (TEST INFRASTRUCTURE: imported by tests/ only).

Follows libs/utils/dynamic_assign.py:83-125 statement by statement in numpy / plain Python (and, through the keyword arguments, its
siblings libs/utils/dynamic_assignV2.py:372-405 and dynamic_assign.py:327-370 `dynamic_k_assign_CF`).  Pinned against the reference
functions themselves, imported from /root/reference by tests/golden/make_dynamic_assign_fixtures.py (fixtures
tests/golden/dynamic_assign_ref.npz): integer outputs, compared exactly.  Ties between equal costs go to the lowest index
(torch.topk leaves them open; the fixtures hold no such tie except complete sets of INFINITY rows).
"""
from __future__ import annotations

import numpy as np

INFINITY = np.float32(987654.0)            # dynamic_assign.py:3


def dynamic_k_assign(cost, pair_wise_ious, n_candidate_k=4, min_k=1, binarize_at=None):
    cost = np.asarray(cost, dtype=np.float32)
    ious = np.array(pair_wise_ious, dtype=np.float32, copy=True)
    num_priors, num_gt = cost.shape
    matching = np.zeros_like(cost)                                            # :95
    if binarize_at is None:
        ious[ious < 0] = 0.0                                                  # :97
    else:
        ious = np.where(ious >= np.float32(binarize_at), np.float32(1), np.float32(0))   # dynamic_k_assign_CF :340-341
    ks = []
    for g in range(num_gt):                                                   # :100-101 topk over the priors, summed, truncated, clamped
        top = np.sort(ious[:, g])[::-1][:n_candidate_k]
        s = np.float32(0)
        for v in top:
            s = np.float32(s + v)
        ks.append(min(max(int(s), min_k), num_priors))
    cost4match = cost.copy()                                                  # :104
    for g in range(num_gt):                                                   # :105-111
        order = np.lexsort((np.arange(num_priors), cost4match[:, g]))         # ascending cost, ties by index
        pos = order[:ks[g]]
        matching[pos, g] = 1.0
        cost4match[pos, :] = INFINITY
    matched = matching.sum(1)                                                 # :114
    multi = matched > 1
    if multi.sum() > 0:                                                       # :116-120
        argmin = np.argmin(cost[multi, :], axis=1)
        matching[multi, :] = 0.0
        matching[np.nonzero(multi)[0], argmin] = 1.0
    prior_idx = np.nonzero(matching.sum(1))[0]                                # :122
    gt_idx = matching[prior_idx].argmax(axis=-1)                              # :123
    return prior_idx.astype(np.int64), gt_idx.astype(np.int64)

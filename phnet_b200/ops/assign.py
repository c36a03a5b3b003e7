"""`dynamic_k_assign` of PHNet's training code (libs/utils/dynamic_assign.py:83-125) on the device: SURVEY.md section 8f row 4.

Same signature and return value as the reference function -- `dynamic_k_assign(cost, pair_wise_ious)` -> (prior_idx, gt_idx),
both int64, priors ascending -- in ONE launch instead of ~10 + 3 per ground truth small torch launches and a host sync per ground
truth (`dynamic_ks[gt_idx].item()`).  The keyword arguments cover the two siblings of the function: `n_candidate_k` / `min_k` are
`max_topk` / `min_topk` of libs/utils/dynamic_assignV2.py:372-405, and `binarize_at=0.8, n_candidate_k=1, min_k=0` is
`dynamic_k_assign_CF` (dynamic_assign.py:327-370).  `dynamic_k_assign_batched` takes [B, num_priors, num_gt] matrices (B images in one
launch, no host sync) and returns padded index tensors and the counts.  Together with `line_iou(..., aligned=False)`, which produces
the IoU matrix, this is the whole of the row; the cost terms around it, autograd and the losses are training and out of scope.
"""
from __future__ import annotations

import torch

from .. import _capi

__all__ = ["dynamic_k_assign", "dynamic_k_assign_batched"]


def dynamic_k_assign_batched(cost: torch.Tensor, pair_wise_ious: torch.Tensor, *, n_candidate_k: int = 4, min_k: int = 1,
                             binarize_at: float | None = None):
    """cost, pair_wise_ious [B, num_priors, num_gt] CUDA.  Returns (prior_idx[B, num_priors], gt_idx[B, num_priors], count[B]), int64:
    image b matched count[b] priors, prior_idx[b, :count[b]] ascending with their ground truths gt_idx[b, :count[b]]."""
    if not (cost.is_cuda and pair_wise_ious.is_cuda) or cost.device != pair_wise_ious.device:
        raise RuntimeError("cost and pair_wise_ious must be CUDA tensors on the same device")
    if cost.dim() != 3 or cost.shape != pair_wise_ious.shape:
        raise RuntimeError("cost and pair_wise_ious must both be [B, num_priors, num_gt]")
    B, P, G = cost.shape
    if G > 0 and P < n_candidate_k:        # torch.topk(ious_matrix, n_candidate_k, dim=0) of the reference raises here
        raise RuntimeError("selected index k out of range")
    c = cost.detach().to(torch.float32).contiguous()
    u = pair_wise_ious.detach().to(torch.float32).contiguous()
    prior_idx = torch.zeros((B, P), dtype=torch.int64, device=c.device)
    gt_idx = torch.zeros((B, P), dtype=torch.int64, device=c.device)
    count = torch.zeros((B,), dtype=torch.int64, device=c.device)
    if P > 0:
        with torch.cuda.device(c.device):
            rc = _capi.lib().phnms_dynamic_k_assign_f32(c.data_ptr(), u.data_ptr(), B, P, G, int(n_candidate_k), int(min_k),
                                                        0 if binarize_at is None else 1,
                                                        0.0 if binarize_at is None else float(binarize_at), prior_idx.data_ptr(),
                                                        gt_idx.data_ptr(), count.data_ptr(),
                                                        torch.cuda.current_stream().cuda_stream)
        _capi.check(rc)
    return prior_idx, gt_idx, count


def dynamic_k_assign(cost: torch.Tensor, pair_wise_ious: torch.Tensor, *, n_candidate_k: int = 4, min_k: int = 1,
                     binarize_at: float | None = None):
    """Drop-in for `libs.utils.dynamic_assign.dynamic_k_assign` (cost, pair_wise_ious [num_priors, num_gt])."""
    if cost.dim() != 2:
        raise RuntimeError("cost must be [num_priors, num_gt]")
    prior_idx, gt_idx, count = dynamic_k_assign_batched(cost[None], pair_wise_ious[None], n_candidate_k=n_candidate_k, min_k=min_k,
                                                        binarize_at=binarize_at)
    n = int(count[0])          # the reference syncs here as well (`nonzero`)
    return prior_idx[0, :n], gt_idx[0, :n]

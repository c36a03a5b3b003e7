"""dynamic_k_assign: the numpy restatement and the CUDA op against outputs of the reference functions
themselves (tests/golden/dynamic_assign_ref.npz: `dynamic_k_assign` and `dynamic_k_assign_CF` imported from /root/reference by
tests/golden/make_dynamic_assign_fixtures.py).  Integer outputs: compared exactly."""
import os

import numpy as np
import pytest

from oracle import dynamic_assign_oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dynamic_assign_ref.npz")


def cases():
    z = np.load(GOLD)
    c = 0
    while f"c{c}_cost" in z:
        cf = bool(z[f"c{c}_cf"][0])
        kw = dict(n_candidate_k=1, min_k=0, binarize_at=0.8) if cf else {}
        yield c, z[f"c{c}_cost"], z[f"c{c}_iou"], kw, z[f"c{c}_prior"], z[f"c{c}_gt"]
        c += 1


def test_fixture_covers_the_branches():
    all_cases = list(cases())
    assert len(all_cases) >= 20 and sum(bool(kw) for _, _, _, kw, _, _ in all_cases) >= 8
    # the last case: priors 3 and 7 are matched by all three columns and keep the ground truth of least cost (2 and 0), not the
    # first column that matched them -- the several-matches branch (dynamic_assign.py:116-120)
    _, _, _, _, prior, gt = all_cases[-1]
    assert prior.tolist() == [3, 7] and gt.tolist() == [2, 0]


def test_oracle_matches_reference_outputs():
    for c, cost, iou, kw, prior, gt in cases():
        p, g = dynamic_assign_oracle.dynamic_k_assign(cost, iou, **kw)
        assert np.array_equal(p, prior) and np.array_equal(g, gt), f"case {c}"


@pytest.mark.gpu
def test_cuda_op_matches_reference_outputs(cuda_device):
    import torch
    from phnet_b200.ops import dynamic_k_assign, dynamic_k_assign_batched
    for c, cost, iou, kw, prior, gt in cases():
        p, g = dynamic_k_assign(torch.from_numpy(cost).to(cuda_device), torch.from_numpy(iou).to(cuda_device), **kw)
        assert p.dtype == torch.int64 and g.dtype == torch.int64
        assert np.array_equal(p.cpu().numpy(), prior) and np.array_equal(g.cpu().numpy(), gt), f"case {c}"
    # many images in one launch: every image equals its own single call (and the oracle)
    rng = np.random.default_rng(5)
    B, P, G = 64, 240, 5
    cost = rng.standard_normal((B, P, G)).astype(np.float32)
    iou = (rng.random((B, P, G)) * 1.5 - 0.3).astype(np.float32)
    pi, gi, cnt = dynamic_k_assign_batched(torch.from_numpy(cost).to(cuda_device), torch.from_numpy(iou).to(cuda_device))
    pi, gi, cnt = pi.cpu().numpy(), gi.cpu().numpy(), cnt.cpu().numpy()
    for b in range(B):
        p, g = dynamic_assign_oracle.dynamic_k_assign(cost[b], iou[b])
        assert cnt[b] == len(p) and np.array_equal(pi[b, :cnt[b]], p) and np.array_equal(gi[b, :cnt[b]], g), f"image {b}"
    # error behaviour of the reference: fewer priors than candidates (torch.topk raises)
    with pytest.raises(RuntimeError):
        dynamic_k_assign(torch.zeros(3, 2, device=cuda_device), torch.zeros(3, 2, device=cuda_device))
    p, g = dynamic_k_assign(torch.zeros(10, 0, device=cuda_device), torch.zeros(10, 0, device=cuda_device))
    assert p.numel() == 0 and g.numel() == 0

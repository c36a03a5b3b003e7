"""line_iou: the numpy restatement and the CUDA op against outputs of the reference function
itself (tests/golden/line_iou_ref.npz).  Floating point: 1e-5 relative (+1e-6 absolute) -- the reference reduces over the
offsets with torch.sum, whose order is not sequential."""
import os

import numpy as np
import pytest

from oracle import line_iou_oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "line_iou_ref.npz")
RTOL, ATOL = 1e-5, 1e-6


def cases():
    z = np.load(GOLD)
    for c in range(5):
        img_w, length = z[f"c{c}_par"]
        yield c, z[f"c{c}_pred"], z[f"c{c}_tgt"], float(img_w), float(length), z[f"c{c}_pair"], z[f"c{c}_aligned"]


def test_oracle_matches_reference_outputs():
    for c, pred, tgt, img_w, length, pair, aligned in cases():
        np.testing.assert_allclose(line_iou_oracle.line_iou(pred, tgt, img_w, length, aligned=False), pair, rtol=RTOL, atol=ATOL)
        n = len(aligned)
        np.testing.assert_allclose(line_iou_oracle.line_iou(pred[:n], tgt[:n], img_w, length, aligned=True), aligned,
                                   rtol=RTOL, atol=ATOL)


@pytest.mark.gpu
def test_cuda_op_matches_reference_outputs(cuda_device):
    import torch
    from phnet_b200.ops import line_iou
    for c, pred, tgt, img_w, length, pair, aligned in cases():
        p, t = torch.from_numpy(pred).to(cuda_device), torch.from_numpy(tgt).to(cuda_device)
        got = line_iou(p, t, img_w, length=length, aligned=False).cpu().numpy()
        np.testing.assert_allclose(got, pair, rtol=RTOL, atol=ATOL, err_msg=f"case {c} pairwise")
        # the op sums sequentially in fp32 exactly like the restatement: bit-exact against it
        assert np.array_equal(got, line_iou_oracle.line_iou(pred, tgt, img_w, length, aligned=False)), f"case {c}"
        n = len(aligned)
        got = line_iou(p[:n], t[:n], img_w, length=length, aligned=True).cpu().numpy()
        np.testing.assert_allclose(got, aligned, rtol=RTOL, atol=ATOL, err_msg=f"case {c} aligned")
    with pytest.raises(RuntimeError):
        line_iou(p[:3], t[:2], img_w, aligned=True)

"""Clip-level get_lanes (prepare -> NMS -> gather) against the reference's own tensor code, line for line, run with torch on
the same GPU (libs/models/Router4OL.py:447-470, RouterV4.py:404-428) and with the reference-equivalent `nms`."""
import pytest
import torch

from phnet_b200.ops import get_lanes, nms
from phnet_b200 import synth

pytestmark = pytest.mark.gpu


def synth_head_output(T, A, n_off, hdr, seed, device):
    """Raw head output: rows (logit0, logit1, start_y, start_x, theta, length, [invalid_len], x...) normalised like the model's."""
    props, _ = synth.make_frames(T, A, n_off, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    out = torch.empty((T, A, hdr + n_off), dtype=torch.float32)
    out[..., 0:2] = torch.randn((T, A, 2), generator=g) * 2.0
    out[..., 2] = props[..., 2]
    out[..., 3] = props[..., 3] / 767.0
    out[..., 4] = torch.rand((T, A), generator=g)
    out[..., 5] = props[..., 4] / (n_off - 1)
    if hdr == 7:
        out[..., 6] = torch.rand((T, A), generator=g) * 0.2
    out[..., hdr:] = props[..., 5:] / 767.0
    return out.to(device)


def reference_get_lanes(predictions, conf_threshold, nms_thres, max_lanes, img_w, n_strips, hdr):
    """One frame, the reference's statements verbatim (with `nms` = the drop-in op)."""
    scores = torch.nn.Softmax(dim=1)(predictions[:, :2])[:, 1]
    keep_inds = scores >= conf_threshold
    predictions = predictions[keep_inds]
    scores = scores[keep_inds]
    if predictions.shape[0] == 0:
        return predictions, keep_inds, None
    nms_predictions = predictions.detach().clone()
    if hdr == 7:
        nms_predictions = torch.cat([nms_predictions[..., :6], nms_predictions[..., 7:]], dim=-1)
    nms_predictions = torch.cat([nms_predictions[..., :4], nms_predictions[..., 5:]], dim=-1)
    nms_predictions[..., 3] = nms_predictions[..., 3] * (img_w - 1)
    nms_predictions[..., 4] = nms_predictions[..., 4] * n_strips
    nms_predictions[..., 5:] = nms_predictions[..., 5:] * (img_w - 1)
    keep, num_to_keep, _ = nms(nms_predictions.contiguous(), scores.contiguous(), overlap=nms_thres, top_k=max_lanes)
    keep = keep[:num_to_keep]
    predictions = predictions[keep]
    if predictions.shape[0]:
        predictions[:, 5] = torch.round(predictions[:, 5] * n_strips)
        if hdr == 7:
            predictions[:, 6] = torch.round(predictions[:, 6] * n_strips)
    return predictions, keep_inds, keep


@pytest.mark.parametrize("hdr,n_off,max_lanes,conf", [(6, 72, 4, 0.5), (6, 36, 4, 0.35), (7, 36, 8, 0.5), (6, 72, 4, 0.999), (6, 72, 4, 0.0)])
def test_get_lanes_matches_reference_statements(cuda_device, hdr, n_off, max_lanes, conf):
    T, A = 24, 240
    out = synth_head_output(T, A, n_off, hdr, seed=hdr * 100 + n_off, device=cuda_device)
    out[3, :, 0:2] = torch.tensor([5.0, -5.0], device=cuda_device)          # a frame where nothing passes the filter
    out[4, 7, 1] = float("nan")                                             # NaN logit: score NaN, filtered out
    lanes, num, index, keep_inds = get_lanes(out, conf, 50, max_lanes, img_w=768)
    torch.cuda.synchronize()
    for t in range(T):
        want, want_mask, keep = reference_get_lanes(out[t].clone(), conf, 50, max_lanes, 768, n_off - 1, hdr)
        assert torch.equal(keep_inds[t], want_mask), f"frame {t}: confidence mask differs"
        assert int(num[t]) == want.shape[0], f"frame {t}: kept {int(num[t])} want {want.shape[0]}"
        n = want.shape[0]
        assert torch.equal(lanes[t, :n], want), f"frame {t}: kept rows differ"
        assert (lanes[t, n:] == 0).all() and (index[t, n:] == 0).all()
        if n:
            orig = torch.nonzero(want_mask).flatten()[keep]
            assert torch.equal(index[t, :n], orig), f"frame {t}: prior indices differ"


def test_scores_are_bitwise_torch_softmax(cuda_device):
    g = torch.Generator().manual_seed(0)
    logits = (torch.randn((64, 240, 2), generator=g) * 6.0).to(cuda_device)
    out = torch.zeros((64, 240, 78), device=cuda_device)
    out[..., :2] = logits
    sm = torch.softmax(logits, dim=2)[..., 1]
    for thr in (0.1, 0.5, 0.9):
        _, _, _, keep_inds = get_lanes(out, thr, 50, 4)
        assert torch.equal(keep_inds, sm >= thr)
    # exact ties with the threshold: pick thresholds equal to actual scores
    thr = float(sm[0, 0])
    _, _, _, keep_inds = get_lanes(out, thr, 50, 4)
    assert torch.equal(keep_inds, sm >= thr)

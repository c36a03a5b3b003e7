"""Clip-level get_lanes (prepare -> NMS -> gather) against the reference's own tensor code, line for line, run with torch on
the same GPU (libs/models/Router4OL.py:447-470, RouterV4.py:404-428) and with the reference-equivalent `nms`."""
import pytest
import torch

from phnet_b200.ops import get_lanes, nms
from phnet_b200 import synth

pytestmark = pytest.mark.gpu


def synth_head_output(T, A, n_off, hdr, seed, device, groups=8):
    return synth.make_head_output(T, A, n_off, hdr, seed, groups=groups).to(device)


UNFUSED = dict(path=1)      # an explicit device path: prepare -> lane NMS -> gather (five launches) instead of the fused kernel


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


@pytest.mark.parametrize("tuning", [None, UNFUSED])
@pytest.mark.parametrize("hdr,n_off,max_lanes,conf,groups", [(6, 72, 4, 0.5, 8), (6, 36, 4, 0.35, 3), (7, 36, 8, 0.5, 4), (6, 72, 4, 0.999, 8),
                                                           (6, 72, 4, 0.0, 2), (6, 72, 8, 0.9, 2), (7, 36, 2, 0.2, 1)])
def test_get_lanes_matches_reference_statements(cuda_device, hdr, n_off, max_lanes, conf, groups, tuning):
    T, A = 24, 240
    out = synth_head_output(T, A, n_off, hdr, seed=hdr * 100 + n_off, device=cuda_device, groups=groups)
    out[3, :, 0:2] = torch.tensor([5.0, -5.0], device=cuda_device)          # a frame where nothing passes the filter
    out[4, 7, 1] = float("nan")                                             # NaN logit: score NaN, filtered out
    out[5, :, 0:2] = torch.tensor([-8.0, 8.0], device=cuda_device)          # every score saturates: all ties
    out[6, 30:, 0:2] = torch.tensor([5.0, -5.0], device=cuda_device)        # <= 32 survivors: torch's unstable small sort
    out[6, :30, 0:2] = torch.round(out[6, :30, 0:2])                        # ... with ties among them
    lanes, num, index, keep_inds = get_lanes(out, conf, 50, max_lanes, img_w=768, tuning=tuning)
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


def test_get_lanes_matches_the_reference_method_bodies(cuda_device):
    """tests/golden/get_lanes_ref.npz: the reference's OWN `get_lanes` (Router4OL.py:437-479, RouterV4.py:394-442), cut out of its
    files with `ast` and executed by tests/golden/make_get_lanes_fixtures.py; inputs are regenerated here from their seeds."""
    import hashlib
    import os
    import numpy as np
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "get_lanes_ref.npz"))
    names = sorted({k.split("/")[0] for k in z.files})
    assert len(names) >= 5
    for nm in names:
        T, A, n_off, hdr, seed, groups, max_lanes = (int(v) for v in z[nm + "/args"])
        conf = float(z[nm + "/conf"][0])
        out = synth.make_head_output(T, A, n_off, hdr, seed=seed, groups=groups, logit_grid=0.25)
        out[3, :, 0:2] = torch.tensor([5.0, -5.0])
        assert hashlib.sha1(out.numpy().tobytes()).digest() == z[nm + "/output_sha1"].tobytes(), f"{nm}: regenerated input differs"
        for tuning in (None, UNFUSED):
            lanes, num, index, keep_inds = get_lanes(out.to(cuda_device), conf, 50, max_lanes, img_w=768, tuning=tuning)
            torch.cuda.synchronize()
            assert np.array_equal(num.cpu().numpy(), z[nm + "/num"]), f"{nm} {tuning}: lanes kept per frame differ"
            assert np.array_equal(lanes.cpu().numpy(), z[nm + "/rows"]), f"{nm} {tuning}: kept rows differ from the reference's"
            for t in range(T):
                assert np.array_equal(keep_inds[t].cpu().numpy(), z[f"{nm}/keep_inds{t}"]), f"{nm} frame {t}: confidence mask differs"


def test_get_lanes_fused_and_unfused_agree_on_a_long_clip(cuda_device):
    for hdr, n_off, K, A in ((6, 72, 4, 240), (7, 36, 8, 240), (6, 36, 4, 1000), (6, 72, 8, 33)):
        out = synth_head_output(700, A, n_off, hdr, seed=A + K, device=cuda_device, groups=3)
        a = get_lanes(out, 0.4, 50, K)
        b = get_lanes(out, 0.4, 50, K, tuning=UNFUSED)
        for x, y in zip(a, b):
            assert torch.equal(x, y), f"hdr={hdr} n_off={n_off} K={K} A={A}"


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

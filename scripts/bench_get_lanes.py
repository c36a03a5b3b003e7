"""Caller-level timing: `get_lanes` + decode for clips of T frames x 240 priors.

reference-style: the reference's statements per frame in a Python loop (softmax, boolean-mask compaction, cat, scalings, nms,
                 keep[:num_to_keep], gather, round; libs/models/Router4OLV2.py:406-448) with (a) the reference's own CUDA op
                 (oracle/_ref, when built) and (b) this repo's drop-in `nms`; decode by the CPU restatement of
                 predictions_to_pred on `.cpu()` rows like the reference (Router4OLV2.py:363-404).
ours:            phnet_b200.ops.get_lanes + decode_lanes for the whole clip, one D2H copy of the points.
Prints one JSON line per configuration (frames/s, wall clock around the whole clip incl. the final host copy).
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import decode_oracle, ref_op  # noqa: E402
from phnet_b200.ops import decode_lanes, get_lanes, nms  # noqa: E402
from tests.test_get_lanes_gpu import synth_head_output  # noqa: E402


def ref_style(out, conf, thres, K, n_off, hdr, nms_fn, decode):
    T = out.shape[0]
    n_strips = n_off - 1
    res = []
    for t in range(T):
        predictions = out[t]
        scores = torch.nn.Softmax(dim=1)(predictions[:, :2])[:, 1]
        keep_inds = scores >= conf
        predictions = predictions[keep_inds]
        scores = scores[keep_inds]
        if predictions.shape[0] == 0:
            res.append([])
            continue
        nms_predictions = predictions.detach().clone()
        if hdr == 7:
            nms_predictions = torch.cat([nms_predictions[..., :6], nms_predictions[..., 7:]], dim=-1)
        nms_predictions = torch.cat([nms_predictions[..., :4], nms_predictions[..., 5:]], dim=-1)
        nms_predictions[..., 3] = nms_predictions[..., 3] * 767
        nms_predictions[..., 4] = nms_predictions[..., 4] * n_strips
        nms_predictions[..., 5:] = nms_predictions[..., 5:] * 767
        keep, num_to_keep, _ = nms_fn(nms_predictions.contiguous(), scores.contiguous(), thres, K)
        keep = keep[:num_to_keep]
        predictions = predictions[keep]
        if predictions.shape[0]:
            predictions[:, 5] = torch.round(predictions[:, 5] * n_strips)
            if hdr == 7:
                predictions[:, 6] = torch.round(predictions[:, 6] * n_strips)
        res.append(decode_oracle.predictions_to_pred(predictions.cpu().numpy(), hdr, 720, 0) if decode else predictions)
    return res


def ours(out, conf, thres, K, decode):
    lanes, num, index, mask = get_lanes(out, conf, thres, K)
    if decode:
        points, npoints, meta = decode_lanes(lanes, num, 720, 0)
        return points.cpu(), npoints.cpu(), meta.cpu()
    return lanes.cpu(), num.cpu()


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def main():
    dev = torch.device("cuda:0")
    for hdr, n_off, K, T in ((6, 72, 4, 16), (6, 72, 4, 256), (7, 36, 8, 100)):
        out = synth_head_output(T, 240, n_off, hdr, seed=1, device=dev)
        rec = {"rows": f"{hdr}+{n_off}", "T": T, "priors": 240, "max_lanes": K}
        for decode in (False, True):
            tag = "+decode" if decode else ""
            if ref_op.load(n_off) is not None:
                rec["ref_loop_ref_op" + tag] = round(T / timed(lambda: ref_style(out, 0.5, 50.0, K, n_off, hdr, ref_op.nms, decode), 3), 1)
            rec["ref_loop_our_nms" + tag] = round(T / timed(lambda: ref_style(out, 0.5, 50.0, K, n_off, hdr,
                                                                          lambda b, s, o, k: nms(b, s, overlap=o, top_k=k), decode), 3), 1)
            rec["ours_clip" + tag] = round(T / timed(lambda: ours(out, 0.5, 50.0, K, decode), 10), 1)
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()

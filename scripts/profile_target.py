"""Small deterministic program for ncu: a few launches of the lane-NMS op on one shape.
env: N NOFF F TOPK GROUPS OUTL REPS and TUNE (a JSON tuning dict, e.g. '{"lanes_per_pass":2}')."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phnet_b200 import _capi, synth  # noqa: E402
from phnet_b200.ops import nms_batched  # noqa: E402

N = int(os.environ.get("N", 1000)); n_off = int(os.environ.get("NOFF", 72)); F = int(os.environ.get("F", 2368))
top_k = int(os.environ.get("TOPK", 4)); reps = int(os.environ.get("REPS", 4))
groups = int(os.environ.get("NGROUPS", os.environ.get("GROUPS", 8))); outl = float(os.environ.get("OUTL", 0.1))
tune = _capi.tuning(**json.loads(os.environ["TUNE"])) if os.environ.get("TUNE") else None
dev = torch.device("cuda:0")
props, scores = synth.make_frames_chunked(F, N, n_off, seed=0, device=dev, groups=groups, outlier_frac=outl)
out = (torch.empty((F, N), dtype=torch.int64, device=dev), torch.empty((F,), dtype=torch.int64, device=dev),
       torch.empty((F, N), dtype=torch.int64, device=dev))
for _ in range(reps):
    nms_batched(props, scores, 50.0, top_k, tuning=tune, out=out)
torch.cuda.synchronize()
print("plan", _capi.plan(F, N, n_off, tune, top_k), "kept", out[1][:8].tolist())

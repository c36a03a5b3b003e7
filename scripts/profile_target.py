"""Small deterministic program for ncu: a few launches of the lane-NMS op on one shape (env: N NOFF F TOPK CLUSTER THREADS PATH REPS)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phnet_b200 import _capi, synth  # noqa: E402
from phnet_b200.ops import nms_batched  # noqa: E402

N = int(os.environ.get("N", 1000)); n_off = int(os.environ.get("NOFF", 72)); F = int(os.environ.get("F", 2368))
top_k = int(os.environ.get("TOPK", 4)); reps = int(os.environ.get("REPS", 4))
tune = _capi.tuning(path=int(os.environ.get("PATH_", 0)), cluster=int(os.environ.get("CLUSTER", 0)),
                    threads=int(os.environ.get("THREADS", 0)))
dev = torch.device("cuda:0")
props, scores = synth.make_frames_chunked(F, N, n_off, seed=0, device=dev)
out = (torch.empty((F, N), dtype=torch.int64, device=dev), torch.empty((F,), dtype=torch.int64, device=dev),
       torch.empty((F, N), dtype=torch.int64, device=dev))
for _ in range(reps):
    nms_batched(props, scores, 50.0, top_k, tuning=tune, out=out)
torch.cuda.synchronize()
print("plan", _capi.plan(F, N, n_off, tune), "kept", out[1][:8].tolist())

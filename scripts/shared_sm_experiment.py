"""What happens when another kernel holds one SM while the persistent NMS kernel runs (e.g. a concurrent NCCL collective):
static frame assignment leaves a second wave, dynamic claiming does not."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phnet_b200 import _capi, synth
from phnet_b200.ops import nms_batched
dev = torch.device("cuda:0")
N, n_off, top_k, F = 1000, 72, 4, 16384
props, scores = synth.make_frames_chunked(F, N, n_off, seed=0, device=dev)
out = (torch.empty((F, N), dtype=torch.int64, device=dev), torch.empty((F,), dtype=torch.int64, device=dev), torch.empty((F, N), dtype=torch.int64, device=dev))
side = torch.cuda.Stream(dev)
res = {}
for name, sched in (("static", 1), ("dynamic", 2)):
    tune = _capi.tuning(path=1, schedule=sched)
    for busy in (False, True):
        for _ in range(3):
            nms_batched(props, scores, 50.0, top_k, tuning=tune, out=out)
        torch.cuda.synchronize()
        if busy:
            with torch.cuda.stream(side):
                torch.cuda._sleep(int(40e6))      # ~20 ms single-thread spin kernel: holds one SM's CTA slot
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            nms_batched(props, scores, 50.0, top_k, tuning=tune, out=out)
        e1.record(); torch.cuda.synchronize()
        res[f"{name}{'_one_sm_busy' if busy else ''}"] = round(F / (e0.elapsed_time(e1) / 5) / 1e3, 3)
print(json.dumps({"Mframes_s": res}))

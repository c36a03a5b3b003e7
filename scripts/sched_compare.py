"""Static vs dynamic frame scheduling, same box, same data (kernel-only frames/s)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phnet_b200 import _capi, synth
from phnet_b200.ops import nms_batched
dev = torch.device("cuda:0")
for N, n_off, top_k, F in [(1000, 72, 4, 16384), (240, 72, 4, 32768), (1000, 36, 4, 16384), (1000, 36, 8, 16384)]:
    props, scores = synth.make_frames_chunked(F, N, n_off, seed=0, device=dev)
    out = (torch.empty((F, N), dtype=torch.int64, device=dev), torch.empty((F,), dtype=torch.int64, device=dev), torch.empty((F, N), dtype=torch.int64, device=dev))
    res = {}
    for rnd in range(2):
        for name, sched in (("static", 1), ("dynamic", 2), ("auto", 0)):
            tune = _capi.tuning(path=1, schedule=sched) if sched else None
            for _ in range(3):
                nms_batched(props, scores, 50.0, top_k, tuning=tune, out=out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                nms_batched(props, scores, 50.0, top_k, tuning=tune, out=out)
            e1.record(); torch.cuda.synchronize()
            res[name] = max(res.get(name, 0), round(F / (e0.elapsed_time(e1) / 10) / 1e3, 3))
    print(json.dumps({"N": N, "n_off": n_off, "top_k": top_k, "Mframes_s": res}), flush=True)

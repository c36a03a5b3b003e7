"""Kernel-only frames/s of the AUTO plan for several shapes (and of the best cluster x threads override)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phnet_b200 import _capi, synth
from phnet_b200.ops import nms_batched

def timeit(props, scores, top_k, tune, out, reps=8):
    for _ in range(3):
        nms_batched(props, scores, 50.0, top_k, tuning=tune, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        nms_batched(props, scores, 50.0, top_k, tuning=tune, out=out)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

dev = torch.device("cuda:0")
shapes = [(1000, 72, 4, 16384), (1000, 72, 8, 8192), (1000, 36, 8, 16384), (1000, 36, 4, 16384), (240, 72, 4, 32768), (240, 36, 8, 32768),
          (2048, 72, 4, 4096), (4096, 72, 4, 2048), (8192, 72, 4, 1024), (256, 72, 4, 32768), (512, 72, 4, 16384)]
if os.environ.get("SHAPES"):
    shapes = [tuple(int(v) for v in s.split("x")) for s in os.environ["SHAPES"].split(",")]
for N, n_off, top_k, F in shapes:
    props, scores = synth.make_frames_chunked(F, N, n_off, seed=0, device=dev)
    out = (torch.empty((F, N), dtype=torch.int64, device=dev), torch.empty((F,), dtype=torch.int64, device=dev),
           torch.empty((F, N), dtype=torch.int64, device=dev))
    bpf = N * (4 * n_off + 40) + 8
    plan = _capi.plan(F, N, n_off, None)
    ms = timeit(props, scores, top_k, None, out)
    rec = {"N": N, "n_off": n_off, "top_k": top_k, "F": F, "auto": {k: plan[k] for k in ("path", "variant", "cluster", "threads", "cols_per_thread")},
           "Mframes_s": round(F / ms / 1e3, 3), "frac": round(F * bpf / ms / 1e6 / 6554.2, 4)}
    if os.environ.get("SEARCH"):
        best = None
        for c in (1, 2, 4, 8, 16):
            for t in (128, 160, 192, 224, 256, 288, 320, 384, 448, 512):
                try:
                    tune = _capi.tuning(path=1, cluster=c, threads=t, variant=2)
                    _capi.plan(F, N, n_off, tune)
                except Exception:
                    continue
                m = timeit(props, scores, top_k, tune, out, reps=4)
                if os.environ.get("SEARCH") == "2":
                    print("   ", c, t, round(F / m / 1e3, 3), flush=True)
                if best is None or m < best[0]:
                    best = (m, c, t)
        rec["best"] = {"cluster": best[1], "threads": best[2], "Mframes_s": round(F / best[0] / 1e3, 3), "frac": round(F * bpf / best[0] / 1e6 / 6554.2, 4)}
    print(json.dumps(rec), flush=True)
    del props, scores, out

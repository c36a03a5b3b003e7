"""Tuning sweep on the GPU box: kernel-only frames/s of the fused path for cluster size x threads (one shape)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phnet_b200 import _capi, synth  # noqa: E402
from phnet_b200.ops import nms_batched  # noqa: E402


def main():
    N = int(os.environ.get("N", 1000))
    n_off = int(os.environ.get("NOFF", 72))
    F = int(os.environ.get("F", 8192))
    top_k = int(os.environ.get("TOPK", 4))
    dev = torch.device("cuda:0")
    props, scores = synth.make_frames_chunked(F, N, n_off, seed=0, device=dev)
    out = (torch.empty((F, N), dtype=torch.int64, device=dev), torch.empty((F,), dtype=torch.int64, device=dev),
           torch.empty((F, N), dtype=torch.int64, device=dev))
    bpf = N * (4 * n_off + 40) + 8
    res = []
    variants = [int(v) for v in os.environ.get('VARIANTS', '2,1').split(',')]
    combos = [(v, c, t, 0) for v in variants for c in (1, 2, 4, 8) for t in (128, 256, 384, 512)]
    for v, c, t, m in combos:
        try:
            tune = _capi.tuning(path=1, cluster=c, threads=t, max_clusters=m, variant=v)
            plan = _capi.plan(F, N, n_off, tune)
        except Exception as e:
            continue
        for _ in range(3):
            nms_batched(props, scores, 50.0, top_k, tuning=tune, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 10
        for _ in range(reps):
            nms_batched(props, scores, 50.0, top_k, tuning=tune, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        r = {"variant": v, "cluster": c, "threads": t, "rows_per_cta": plan["rows_per_cta"], "smem": plan["smem_bytes"], "grid": plan["grid"], "occ": plan["max_active_clusters"],
             "ms": round(ms, 4), "Mframes_s": round(F / ms / 1e3, 3), "GBs": round(F * bpf / ms / 1e6, 1),
             "frac": round(F * bpf / ms / 1e6 / 6554.2, 4)}
        res.append(r)
        print(json.dumps(r), flush=True)
    best = max(res, key=lambda r: r["Mframes_s"])
    print("BEST", json.dumps(best))


if __name__ == "__main__":
    main()

"""Per-call latency of the drop-in op, the way PHNet calls it: ONE frame per call (libs/models/Router4OL.py:460-465).
Prints calls/s with and without the `keep[:num_to_keep]` host sync, for this repo's op and -- when oracle/_ref is built -- for the
reference's own CUDA op on the same GPU; and GPU time per call (CUDA events over a burst of async calls)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_op  # noqa: E402
from phnet_b200 import synth  # noqa: E402
from phnet_b200.ops import nms  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    out = {}
    F = 512
    for n_off, N, top_k in ((72, 240, 4), (36, 240, 8), (72, 1000, 4)):
        props, scores = synth.make_frames(F, N, n_off, seed=3, groups=4)
        props, scores = props.to(dev), scores.to(dev)
        impls = [("ours", lambda b, s, o, k: nms(b, s, overlap=o, top_k=k))]
        if ref_op.path(n_off) is not None:
            impls.append(("reference", ref_op.nms))
        for name, fn in impls:
            for sync in (True, False):
                for f in range(16):
                    fn(props[f], scores[f], 50.0, top_k)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0 = time.perf_counter()
                e0.record()
                for f in range(F):
                    keep, num, _ = fn(props[f], scores[f], 50.0, top_k)
                    if sync:
                        keep = keep[:num]
                e1.record()
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                key = f"{name}_N{N}_No{n_off}_k{top_k}_{'sync' if sync else 'async'}"
                out[key + "_calls_per_s"] = round(F / dt, 1)
                if not sync:
                    out[key + "_gpu_us_per_call"] = round(e0.elapsed_time(e1) * 1e3 / F, 2)
            # the same async burst over frames sliced beforehand (get_lanes hands over freshly filtered tensors, not views it
            # indexes per call): the op call alone
            ps, ss = [props[f] for f in range(F)], [scores[f] for f in range(F)]
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for f in range(F):
                fn(ps[f], ss[f], 50.0, top_k)
            torch.cuda.synchronize()
            out[f"{name}_N{N}_No{n_off}_k{top_k}_async_presliced_calls_per_s"] = round(F / (time.perf_counter() - t0), 1)
            if name == "ours":      # CUDA-graph replay for the fixed shape: with the two input copies, and the replay alone
                from phnet_b200.ops import GraphedNMS
                g = GraphedNMS(N, n_off, 50.0, top_k, device=dev)
                for mode in ("copy+replay", "replay"):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    for f in range(F):
                        if mode == "replay":
                            g.replay()
                        else:
                            g(ps[f], ss[f])
                    torch.cuda.synchronize()
                    out[f"graph_N{N}_No{n_off}_k{top_k}_{mode}_calls_per_s"] = round(F / (time.perf_counter() - t0), 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()

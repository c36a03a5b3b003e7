"""Times the REFERENCE CUDA op (oracle/_ref, built from /root/reference/libs/ops/csrc by oracle/build_ref.py) on the
B200 the way get_lanes drives it: one call per frame followed by `keep[:num_to_keep]` (a host sync,
libs/models/Router4OL.py:460-465), and once more without the sync.  Test infrastructure: not part of bench.py."""
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
    for n_off, N in ((72, 1000), (72, 240), (36, 240)):
        if ref_op.load(n_off) is None:
            continue
        F = 256
        props, scores = synth.make_frames(F, N, n_off, seed=3)
        props, scores = props.to(dev), scores.to(dev)
        for name, fn in (("reference", ref_op.nms), ("ours", lambda b, s, o, k: nms(b, s, overlap=o, top_k=k))):
            for sync in (True, False):
                for f in range(8):
                    fn(props[f], scores[f], 50.0, 4)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for f in range(F):
                    keep, num, _ = fn(props[f], scores[f], 50.0, 4)
                    if sync:
                        keep = keep[:num]
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                out[f"{name}_N{N}_No{n_off}_{'sync' if sync else 'async'}_frames_per_s"] = round(F / dt, 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()

"""Small program for compute-sanitizer (memcheck / racecheck / initcheck / synccheck): every device path once, tiny batches.

    compute-sanitizer --tool memcheck  python scripts/sanitize_target.py
    compute-sanitizer --tool racecheck python scripts/sanitize_target.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phnet_b200 import peer, synth  # noqa: E402
from phnet_b200.ops import decode_lanes, get_lanes, nms, nms_batched  # noqa: E402

dev = torch.device("cuda:0")
runs = 0
for N, n_off, F in ((1000, 72, 6), (1000, 36, 6), (240, 72, 5), (33, 36, 3), (20, 72, 3)):
    props, scores = synth.make_frames(F, N, n_off, seed=N + n_off, ties=(N < 100))
    p, s = props.to(dev), scores.to(dev)
    for top_k in (4, 0, 64):                       # 0 / 64: many fallback batches (cluster exchange, spare lanes)
        for tuning in (None, dict(path=1, cluster=2), dict(path=1, cluster=4, threads=256), dict(path=1, variant=1), dict(path=2),
                       dict(path=1, schedule=2), dict(path=1, schedule=1)):
            try:
                nms_batched(p, s, 50.0, top_k, tuning=tuning)
                runs += 1
            except Exception as e:                 # an override that does not fit the shape
                if "tuning" not in str(e):
                    raise
    buf = [torch.zeros((F + 2, 5), dtype=torch.int64, device=dev) for _ in range(2)]
    nms_batched(p, s, 50.0, 4, collect=peer.local_collect(buf, 1))
    nms(p[0], s[0], overlap=50, top_k=4)
for hdr, n_off in ((6, 72), (7, 36)):
    out = torch.rand((5, 240, hdr + n_off), device=dev)
    out[..., :2] = torch.randn((5, 240, 2), device=dev) * 2
    lanes, num, index, mask = get_lanes(out, 0.4, 50.0, 4 if hdr == 6 else 8)
    decode_lanes(lanes, num, 720, 100)
torch.cuda.synchronize()
print("sanitize target ok:", runs, "nms_batched calls")

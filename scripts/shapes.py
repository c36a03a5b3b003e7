"""Kernel-only frames/s and HBM-roofline fraction of the AUTO plan over shapes AND input distributions.

    python scripts/shapes.py                       # the default table (profiles/r2_shapes.log)
    SHAPES=1000x72x4x16384x2x0.1,... python scripts/shapes.py     # N x n_off x top_k x F x groups x outlier_frac
    TUNE='{"variant":2,"path":1}' python scripts/shapes.py        # a tuning override for every line

`groups` = lane groups per frame of the generator (phnet_b200/synth.py): 8 is the bench default, 2-4 is what roads look like
(PHNet max_lanes = 4); outlier_frac = 0 means the frame has fewer lanes than top_k and the greedy scan runs to the end.
"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phnet_b200 import _capi, synth
from phnet_b200.ops import nms_batched

PEAK = 6554.2
try:
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as fh:
        PEAK = float(json.load(fh)["hbm_gbs"])
except Exception:
    pass


def timeit(props, scores, top_k, tune, out, reps=8):
    for _ in range(3):
        nms_batched(props, scores, 50.0, top_k, tuning=tune, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        nms_batched(props, scores, 50.0, top_k, tuning=tune, out=out)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


dev = torch.device("cuda:0")
shapes = [
    # headline shape, generator default and road-like frames
    (1000, 72, 4, 16384, 8, 0.1), (1000, 72, 4, 16384, 4, 0.1), (1000, 72, 4, 16384, 3, 0.1), (1000, 72, 4, 16384, 2, 0.1),
    (1000, 72, 4, 16384, 1, 0.1), (1000, 72, 4, 16384, 2, 0.0), (1000, 72, 4, 16384, 3, 0.01),
    # top_k = 8 (VIL-100: optionsV3.py:89-91)
    (1000, 72, 8, 16384, 8, 0.1), (1000, 36, 8, 16384, 8, 0.1), (240, 36, 8, 32768, 8, 0.1), (240, 36, 8, 32768, 3, 0.1),
    (1000, 36, 4, 16384, 8, 0.1), (240, 72, 4, 32768, 8, 0.1), (240, 72, 4, 32768, 3, 0.1), (240, 36, 4, 32768, 3, 0.1),
    (100, 72, 4, 65536, 4, 0.1),
    # stress sweep sizes (BASELINE config 4)
    (256, 72, 4, 32768, 8, 0.1), (512, 72, 4, 16384, 8, 0.1), (2048, 72, 4, 4096, 8, 0.1), (4096, 72, 4, 2048, 8, 0.1),
    (8192, 72, 4, 1024, 8, 0.1),
]
if os.environ.get("SHAPES"):
    shapes = []
    for s in os.environ["SHAPES"].split(","):
        v = s.split("x")
        shapes.append((int(v[0]), int(v[1]), int(v[2]), int(v[3]), int(v[4]) if len(v) > 4 else 8, float(v[5]) if len(v) > 5 else 0.1))
tunes = [None]
if os.environ.get("TUNE"):
    tunes = [json.loads(os.environ["TUNE"])]
if os.environ.get("TUNES"):      # a JSON list of tuning dicts (null = auto): every shape is timed under each
    tunes = json.loads(os.environ["TUNES"])
for N, n_off, top_k, F, groups, outl in shapes:
    props, scores = synth.make_frames_chunked(F, N, n_off, seed=0, device=dev, groups=groups, outlier_frac=outl)
    out = (torch.empty((F, N), dtype=torch.int64, device=dev), torch.empty((F,), dtype=torch.int64, device=dev),
           torch.empty((F, N), dtype=torch.int64, device=dev))
    bpf = N * (4 * n_off + 40) + 8
    for td in tunes:
        tune = _capi.tuning(**td) if td else None
        try:
            plan = _capi.plan(F, N, n_off, tune, top_k)
            med, best = timeit(props, scores, top_k, tune, out)
        except _capi.PhnmsError as e:
            print(json.dumps({"N": N, "n_off": n_off, "top_k": top_k, "tune": td, "error": str(e)}), flush=True)
            continue
        rec = {"N": N, "n_off": n_off, "top_k": top_k, "F": F, "groups": groups, "outliers": outl,
               "plan": {k: plan[k] for k in ("path", "variant", "cluster", "threads", "grid")},
               "ms": round(med, 4), "Mframes_s": round(F / med / 1e3, 3), "frac": round(F * bpf / med / 1e6 / PEAK, 4),
               "frac_best": round(F * bpf / best / 1e6 / PEAK, 4), "mean_kept": round(float(out[1].float().mean()), 3)}
        if td:
            rec["tune"] = td
        print(json.dumps(rec), flush=True)
    del props, scores, out

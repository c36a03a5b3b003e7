"""Soak test: many back-to-back launches per shape; every launch must reproduce the first one bit for bit, and a sample
of frames must match the CPU oracle.  Guards against timing-dependent races (two were found and fixed in round 1)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phnet_b200 import _capi, synth
from phnet_b200.ops import nms_batched
from tests.util import assert_same, oracle_batched

dev = torch.device("cuda:0")
reps = int(os.environ.get("REPS", 30))
shapes = [(1000, 72, 4, 8192, None), (1000, 72, 8, 4096, None), (1000, 72, 0, 1024, None), (1000, 36, 8, 8192, None), (240, 72, 4, 16384, None),
          (240, 36, 8, 16384, None), (2048, 72, 4, 2048, None), (4096, 72, 4, 1024, None), (8192, 72, 4, 512, None), (700, 72, 4, 8192, dict(path=1, cluster=4, variant=2)),
          (1000, 72, 4, 4096, dict(path=1, variant=1)), (300, 50, 4, 4096, None)]
for N, n_off, top_k, F, tune in shapes:
    props, scores = synth.make_frames_chunked(F, N, n_off, seed=N + n_off, device=dev, ties=(N == 240))
    g = torch.Generator().manual_seed(N)
    nv = torch.randint(N // 2, N + 1, (F,), generator=g, dtype=torch.int32).to(dev) if N in (240, 700) else None
    first = None
    t0 = time.time()
    for r in range(reps):
        out = nms_batched(props, scores, 50.0, top_k, nv, tuning=tune)
        if r % 5 == 0:
            torch.cuda.synchronize()
        if first is None:
            torch.cuda.synchronize()
            first = [t.clone() for t in out]
            idx = torch.arange(0, F, max(1, F // 24))[:24]
            want = oracle_batched(props[idx].cpu(), scores[idx].cpu(), 50.0, top_k, None if nv is None else nv[idx].cpu())
            assert_same([t[idx] for t in out], want, f"N={N} No={n_off} top_k={top_k}")
        else:
            assert all(torch.equal(a, b) for a, b in zip(out, first)), f"launch {r} differs: N={N} No={n_off} top_k={top_k}"
    torch.cuda.synchronize()
    print(f"ok N={N} No={n_off} top_k={top_k} F={F} tune={tune} plan={ {k: v for k, v in _capi.plan(F, N, n_off, _capi.tuning(**tune) if tune else None).items() if k in ('path','variant','cluster','threads','cols_per_thread')} } {reps} launches {time.time() - t0:.1f}s", flush=True)
print("SOAK OK")

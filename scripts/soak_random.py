"""Randomised soak: ITER (default 10 000) launches over random configurations of the lane-NMS op -- shape, offset count, top_k,
threshold, generator (lane groups, outliers, ties), ragged n_valid, device algorithm (auto / streaming with random warps, lanes
per pass and draw cap / the one-launch small-frame kernel with random grids / cluster kernels with random cluster size and schedule / tiled), stream -- each checked bit for bit
against the CPU oracle on a sample of its frames, and every configuration launched three times in a row with identical results.
Every third configuration also collects the compact kept-lane records and checks them.
A hang becomes a failed launch (mbarrier waits are bounded, common.cuh), a race a mismatch.  Prints one line per 500 launches.

    ITER=10000 SEED=0 python scripts/soak_random.py
"""
import os, random, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phnet_b200 import _capi, peer, sharding, synth
from phnet_b200.ops import nms_batched
from tests.util import assert_same, oracle_batched

dev = torch.device("cuda:0")
iters = int(os.environ.get("ITER", 10000))
rng = random.Random(int(os.environ.get("SEED", 0)))
side = torch.cuda.Stream()
t0 = time.time()
launches = configs = 0
cache = {}
while launches < iters:
    n_off = rng.choice([72, 72, 36])
    N = rng.choice([1, 7, 31, 32, 33, 64, 100, 240, 240, 256, 333, 500, 1000, 1000, 1000, 1024, 1500, 2048, 4096])
    F = rng.choice([1, 2, 5, 37, 148, 149, 300, 1000, 3000]) if N <= 1024 else rng.choice([1, 3, 40, 200])
    while F * N > 1_500_000:
        F //= 2
    top_k = rng.choice([1, 2, 3, 4, 4, 4, 5, 8, 8, 0, N])
    thr = rng.choice([10.0, 20.0, 30.0, 40.0, 50.0, 50.0])
    groups, outl, ties = rng.choice([1, 2, 3, 4, 8]), rng.choice([0.0, 0.01, 0.1]), rng.random() < 0.2
    kind = rng.choice(["auto", "auto", "stream", "stream", "cluster", "tiled", "small"])
    tune = None
    if kind == "stream" and 1 <= top_k <= 8:
        tune = dict(variant=3, stream_warps=rng.choice([0, 1, 3, 8, 16]), lanes_per_pass=rng.choice([0, 1, 2, 4]),
                    select_cap=rng.choice([0, 8, 16, 64, 256]))
    elif kind == "cluster":
        tune = dict(path=1, variant=rng.choice([1, 2]), cluster=rng.choice([0, 0, 1, 2, 4, 8]), schedule=rng.choice([0, 1, 2]))
    elif kind == "tiled":
        tune = dict(path=2)
    elif kind == "small" and N <= 512:
        tune = dict(variant=4, max_clusters=rng.choice([0, 0, 1, 5, 148]))
    if tune is not None:
        try:
            _capi.plan(F, N, n_off, _capi.tuning(**tune), top_k)
        except _capi.PhnmsError:
            continue      # this override does not fit the shape
    key = (F, N, n_off, groups, outl, ties)
    if key not in cache:
        if len(cache) > 24:
            cache.clear()
        cache[key] = synth.make_frames_chunked(F, N, n_off, seed=rng.randrange(1 << 30), device=dev, groups=groups, outlier_frac=outl, ties=ties)
    props, scores = cache[key]
    nv = None
    if rng.random() < 0.3:
        nv = torch.randint(0, N + 1, (F,), generator=torch.Generator().manual_seed(rng.randrange(1 << 30)), dtype=torch.int32).to(dev)
    ctx = f"F={F} N={N} No={n_off} top_k={top_k} thr={thr} groups={groups} outl={outl} ties={ties} ragged={nv is not None} tune={tune}"
    use_side = rng.random() < 0.3
    # every third configuration also collects the compact kept-lane records (stored by the NMS kernels themselves on the
    # streaming / small-frame / cluster paths, by the record kernel elsewhere) into two local buffers at a random row offset
    bufs, row0 = None, 0
    if top_k >= 1 and top_k <= 64 and rng.random() < 0.33:
        row0 = rng.choice([0, 3])
        bufs = [torch.full((F + row0 + 2, top_k + 1), -7, dtype=torch.int64, device=dev) for _ in range(2)]
    outs = []
    try:
        for _ in range(3):
            if use_side:
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    outs.append(nms_batched(props, scores, thr, top_k, nv, tuning=tune, collect=None if bufs is None else peer.local_collect(bufs, row0)))
                torch.cuda.current_stream().wait_stream(side)
            else:
                outs.append(nms_batched(props, scores, thr, top_k, nv, tuning=tune, collect=None if bufs is None else peer.local_collect(bufs, row0)))
        torch.cuda.synchronize()
    except Exception as e:   # noqa: BLE001
        print("FAILED LAUNCH:", ctx, repr(e), flush=True)
        raise
    for o in outs[1:]:
        assert all(torch.equal(a, b) for a, b in zip(o, outs[0])), "repeat differs: " + ctx
    if bufs is not None:
        want_rec = sharding.pack_kept(outs[0][0], outs[0][1], top_k)
        for b in bufs:
            assert torch.equal(b[row0:row0 + F], want_rec), "records differ: " + ctx
            assert bool((b[:row0] == -7).all()) and bool((b[row0 + F:] == -7).all()), "records outside the call's rows: " + ctx
    idx = torch.arange(0, F, max(1, F // 12))[:12]
    want = oracle_batched(props[idx].cpu(), scores[idx].cpu(), thr, top_k, None if nv is None else nv[idx].cpu())
    assert_same([t[idx] for t in outs[0]], want, ctx)
    launches += 3
    configs += 1
    if launches % 498 < 3:
        print(f"{launches} launches, {configs} configurations ok, {time.time() - t0:.0f} s", flush=True)
print(f"SOAK OK: {launches} launches over {configs} random configurations in {time.time() - t0:.0f} s")

"""GPU parity tests of the one-launch small-call kernel (phnet_b200/csrc/small.cuh: one CTA per frame, rows in registers, one
greedy round per kept lane) against the CPU oracle, bit-exact on keep / num / parent: the call PHNet itself makes (ONE frame of
<= 240 priors, libs/models/Router4OL.py:460-465) and small batches; every top_k incl. 0 and N, edge frames, ties under all sort
models (the bitonic replay for <= 32 proposals), ragged and misaligned inputs, and -- forced by tuning -- thousands of frames."""
import pytest
import torch

from phnet_b200 import _capi, synth
from phnet_b200.ops import nms, nms_batched, plan
from tests.util import assert_same, oracle_batched

pytestmark = pytest.mark.gpu

SMALL = dict(variant=_capi.FUSED_SMALL)


def run_both(props, scores, thr, top_k, dev, n_valid=None, tuning=None, sort_model=0, ctx=""):
    nv = None if n_valid is None else n_valid.to(dev)
    got = nms_batched(props.to(dev), scores.to(dev), thr, top_k, nv, tuning=tuning, sort_model=sort_model)
    torch.cuda.synchronize()
    assert_same(got, oracle_batched(props, scores, thr, top_k, n_valid, sort_model=sort_model), ctx)
    return got


def test_plan_small_calls_take_the_one_launch_kernel(cuda_device):
    for n_off in (36, 72):
        for F, N in ((1, 240), (1, 1), (1, 512), (8, 240), (4, 512), (64, 32)):
            for top_k in (0, 1, 4, 8, N):
                pl = plan(F, N, n_off, top_k=top_k)
                assert pl["variant"] == _capi.FUSED_SMALL and pl["launches"] == 1 and pl["workspace_bytes"] == 0 and pl["grid"] == F, pl
        assert plan(1, 513, n_off, top_k=4)["variant"] != _capi.FUSED_SMALL          # more than 512 proposals per frame
        assert plan(16, 300, n_off, top_k=4)["variant"] == _capi.FUSED_STREAM        # more than 2048 proposals in the call, N > 256
        big = plan(100000, 240, n_off, top_k=4)     # frames of <= 256 proposals: this kernel whatever the batch size, persistent CTAs
        assert big["variant"] == _capi.FUSED_SMALL and big["launches"] == 1 and 148 <= big["grid"] <= 148 * 8
    assert plan(1, 240, 50, top_k=4)["variant"] == _capi.FUSED_SMEM                  # other offset counts: the shared-memory kernel
    with pytest.raises(_capi.PhnmsError):
        plan(4, 600, 72, tuning=SMALL)


@pytest.mark.parametrize("n_off", [72, 36])
@pytest.mark.parametrize("N", [1, 2, 5, 31, 32, 33, 64, 65, 100, 240, 256, 333, 500, 512])
def test_one_frame_calls(cuda_device, N, n_off):
    """The drop-in call: nms(boxes[N, 5+No], scores[N], overlap, top_k)."""
    for groups, outl in ((8, 0.1), (3, 0.0), (1, 0.02)):
        props, scores = synth.make_frames(3, N, n_off, seed=N * 7 + groups, groups=min(groups, max(1, N // 4)), outlier_frac=outl)
        for top_k in (0, 1, 2, 4, 8, N, N + 5):
            want = oracle_batched(props, scores, 50.0, top_k)
            for f in range(3):
                keep, num, parent = nms(props[f].to(cuda_device), scores[f].to(cuda_device), 50.0, top_k)
                assert num.dim() == 0
                assert_same((keep[None], num.reshape(1), parent[None]), [w[f:f + 1] for w in want],
                            f"one frame N={N} No={n_off} top_k={top_k} groups={groups} f={f}")


@pytest.mark.parametrize("n_off", [72, 36])
def test_edge_frames_ties_and_sort_models(cuda_device, n_off):
    for seed in range(8):
        p, s = synth.edge_frame(n_off, seed=seed)
        for top_k in (0, 1, 4, 8, 96):
            for thr in (50.0, 0.0, -1.0, float("nan"), float("inf")):
                run_both(p[None], s[None], thr, top_k, cuda_device, ctx=f"edge seed={seed} No={n_off} top_k={top_k} thr={thr}")
    for N in (5, 20, 32, 33, 100, 500):
        props, scores = synth.make_frames(4, N, n_off, seed=N, ties=True)
        scores[0, : N // 2] = 1.0
        if N > 4:
            scores[1, 1] = float("nan")
            scores[1, 3] = -float("nan")
            scores[2, ::2] = 0.0
            scores[2, 1::4] = -0.0
        for sm in (0, 1, 2):
            for top_k in (4, 0):
                run_both(props, scores, 50.0, top_k, cuda_device, sort_model=sm, ctx=f"ties N={N} No={n_off} sort_model={sm} top_k={top_k}")


def test_ragged_and_misaligned(cuda_device):
    F, N = 6, 333            # 333 * 77 words: frames start at every alignment modulo 16 bytes
    for n_off in (72, 36):
        P = 5 + n_off
        props, scores = synth.make_frames(F, N, n_off, seed=8, groups=3)
        n_valid = torch.tensor([0, 1, N, 32, 20, 33], dtype=torch.int32)
        run_both(props, scores, 50.0, 4, cuda_device, n_valid=n_valid, ctx=f"ragged No={n_off}")
        big = torch.zeros(F * N * P + 3, device=cuda_device)
        sbig = torch.zeros(F * N + 3, device=cuda_device)
        for shift in (0, 1, 2, 3):
            view = big[shift: shift + F * N * P].view(F, N, P)
            view.copy_(props)
            sview = sbig[(3 - shift): (3 - shift) + F * N].view(F, N)
            sview.copy_(scores)
            st = torch.cuda.Stream()
            st.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(st):
                got = nms_batched(view, sview, 50.0, 4, n_valid.to(cuda_device))
                one = nms(view[2], sview[2], 50.0, 4)            # a frame in the middle of a tensor: 4-byte aligned only
            st.synchronize()
            want = oracle_batched(props, scores, 50.0, 4, n_valid)
            assert_same(got, want, f"shift={shift} No={n_off}")
            assert_same((one[0][None], one[1].reshape(1), one[2][None]), [w[2:3] for w in want], f"one frame, shift={shift} No={n_off}")


@pytest.mark.parametrize("N,n_off,top_k", [(240, 72, 4), (240, 36, 8), (512, 72, 4), (100, 36, 0), (33, 72, 33), (64, 72, 8)])
def test_forced_on_big_batches_agrees_with_the_streaming_path_and_repeats(cuda_device, N, n_off, top_k):
    F = 3000
    props, scores = synth.make_frames_chunked(F, N, n_off, seed=N + top_k, device=cuda_device, groups=3)
    other = dict(variant=_capi.FUSED_STREAM) if 1 <= top_k <= 8 else dict(path=1, variant=_capi.FUSED_REG)
    ref = nms_batched(props, scores, 50.0, top_k, tuning=other)     # the streaming / cluster kernels
    for rep in range(3):
        got = nms_batched(props, scores, 50.0, top_k, tuning=SMALL if rep else dict(variant=_capi.FUSED_SMALL, max_clusters=5))
        for x, y in zip(got, ref):
            assert torch.equal(x, y), f"N={N} No={n_off} top_k={top_k} rep={rep}"
    idx = torch.arange(0, F, 131)
    assert_same([t[idx] for t in ref], oracle_batched(props[idx].cpu(), scores[idx].cpu(), 50.0, top_k), f"oracle sample N={N}")


@pytest.mark.parametrize("n_off", [72, 36])
def test_persistent_ctas_over_ragged_frames(cuda_device, n_off):
    """Many frames per CTA (max_clusters caps the grid), every frame with its own number of real proposals -- warps whose slice of a
    frame is empty skip that frame's load, so the per-warp barrier phases advance at different rates -- on a side stream, from
    tensors that are only 4-byte aligned."""
    F, N = 2000, 250
    g = torch.Generator().manual_seed(n_off)
    props, scores = synth.make_frames(F, N, n_off, seed=21, groups=3)
    n_valid = torch.randint(0, N + 1, (F,), generator=g, dtype=torch.int32)
    n_valid[:8] = torch.tensor([0, 0, 1, N, 32, 33, 0, 64], dtype=torch.int32)
    P = 5 + n_off
    big = torch.zeros(F * N * P + 1, device=cuda_device)[1:].view(F, N, P)
    big.copy_(props)
    want = oracle_batched(props, scores, 50.0, 4, n_valid)
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        for cap in (0, 1, 7, 148):
            got = nms_batched(big, scores.to(cuda_device), 50.0, 4, n_valid.to(cuda_device), tuning=dict(variant=_capi.FUSED_SMALL, max_clusters=cap))
            st.synchronize()
            assert_same(got, want, f"ragged persistent No={n_off} max_clusters={cap}")


def test_subnormal_offsets_and_thresholds(cuda_device):
    for n_off in (72, 36):
        props, scores = synth.make_frames(4, 300, n_off, seed=11)
        props[..., 5:] *= 1e-41
        for thr in (50e-41, 5e-41, 1e-45):
            run_both(props, scores, thr, 4, cuda_device, ctx=f"subnormal No={n_off} thr={thr}")
        props, scores = synth.make_frames(6, 240, n_off, seed=12, groups=3)
        for thr in (10.0, 20.0, 30.0, 40.0, 50.0):
            run_both(props, scores, thr, 4, cuda_device, ctx=f"No={n_off} thr={thr}")


def test_graph_replay_of_the_one_frame_call(cuda_device):
    """GraphedNMS: the captured launch, replayed on new contents of its static buffers, equals the eager call -- fixed-size frames,
    ragged frames (n_valid inside the graph), and a frame of more than 512 proposals (the cluster kernel with its workspace)."""
    from phnet_b200.ops import GraphedNMS
    for N, n_off, top_k in ((240, 72, 4), (240, 36, 8), (1000, 72, 4)):
        props, scores = synth.make_frames(6, N, n_off, seed=N + top_k, groups=3)
        p, s = props.to(cuda_device), scores.to(cuda_device)
        g = GraphedNMS(N, n_off, 50.0, top_k, device=cuda_device)
        for f in range(6):
            keep, num, parent = g(p[f], s[f])
            want = nms(p[f], s[f], 50.0, top_k)
            assert torch.equal(keep, want[0]) and int(num) == int(want[1]) and torch.equal(parent, want[2]), f"N={N} f={f}"
        g.boxes.copy_(p[0]); g.scores.copy_(s[0])             # the caller filled the static buffers itself
        keep, num, parent = g.replay()
        assert torch.equal(keep, nms(p[0], s[0], 50.0, top_k)[0])
        if N <= 512:
            r = GraphedNMS(N, n_off, 50.0, top_k, device=cuda_device, ragged=True)
            for n in (N, 1, 33, 100, 0):
                keep, num, parent = r(p[1, :n].contiguous(), s[1, :n].contiguous())
                want = nms(p[1, :n].contiguous(), s[1, :n].contiguous(), 50.0, top_k)
                assert int(num) == int(want[1]) and torch.equal(keep[:n], want[0]) and torch.equal(parent[:n], want[2]), f"ragged n={n}"
                assert bool((keep[n:] == 0).all()) and bool((parent[n:] == 0).all())

"""GPU parity tests of the streaming fast path (select -> stream -> resume; phnet_b200/csrc/select.cuh, stream.cuh) against the
CPU oracle, bit-exact on keep / num / parent.  The inputs are chosen to drive every branch of the path: frames whose kept
lanes sit deep in the order (few lane groups: the draw loop runs long), frames with fewer lanes than top_k (the scan
exhausts the frame), draw caps small enough that frames stay open and are redone by the resume pass, every lanes-per-pass /
warps-per-CTA / ring configuration, fewer frames than SMs (frames cut into units), ragged and misaligned batches."""
import pytest
import torch

from phnet_b200 import _capi, synth
from phnet_b200.ops import nms_batched, plan
from tests.util import assert_same, oracle_batched

pytestmark = pytest.mark.gpu

STREAM = dict(variant=_capi.FUSED_STREAM)


def run_both(props, scores, thr, top_k, dev, n_valid=None, tuning=None, sort_model=0, ctx=""):
    nv = None if n_valid is None else n_valid.to(dev)
    got = nms_batched(props.to(dev), scores.to(dev), thr, top_k, nv, tuning=tuning, sort_model=sort_model)
    torch.cuda.synchronize()
    want = oracle_batched(props, scores, thr, top_k, n_valid, sort_model=sort_model)
    assert_same(got, want, ctx)
    return got


def test_default_plan_is_the_streaming_path(cuda_device):
    for n_off in (36, 72):
        for top_k in (1, 4, 8):
            assert plan(64, 1000, n_off, top_k=top_k)["variant"] == _capi.FUSED_STREAM
        assert plan(64, 1000, n_off, top_k=0)["variant"] == _capi.FUSED_REG
        assert plan(64, 1000, n_off, top_k=9)["variant"] == _capi.FUSED_REG


@pytest.mark.parametrize("n_off", [72, 36])
@pytest.mark.parametrize("groups", [1, 2, 3, 4, 8])
def test_lane_group_counts(cuda_device, groups, n_off):
    """2-4 lanes per frame is what roads look like (PHNet max_lanes = 4): the kept lanes beyond the groups come from deep in
    the order (outliers) or do not exist at all (outlier_frac = 0: the scan exhausts the frame and keeps < top_k)."""
    for N in (240, 1000):
        for outl in (0.1, 0.0, 0.01):
            props, scores = synth.make_frames(12, N, n_off, seed=groups * 31 + N, groups=groups, outlier_frac=outl)
            for top_k in (4, 8):
                run_both(props, scores, 50.0, top_k, cuda_device, tuning=STREAM,
                         ctx=f"groups={groups} N={N} No={n_off} outliers={outl} top_k={top_k}")


@pytest.mark.parametrize("cap", [8, 16, 24, 64, 4096])
def test_draw_cap_and_resume_pass(cuda_device, cap):
    """A small cap leaves frames open; those that still have an uncovered proposal after the streaming pass are redone by the
    register-resident cluster kernel through the device-side resume list.  Results never depend on the cap."""
    for n_off, N, groups, outl in ((72, 1000, 2, 0.1), (36, 1000, 1, 0.05), (72, 300, 3, 0.0), (72, 2048, 2, 0.02)):
        props, scores = synth.make_frames(20, N, n_off, seed=cap + N, groups=groups, outlier_frac=outl)
        for top_k in (1, 3, 4, 8):
            run_both(props, scores, 50.0, top_k, cuda_device, tuning=dict(variant=3, select_cap=cap),
                     ctx=f"cap={cap} N={N} No={n_off} groups={groups} top_k={top_k}")


@pytest.mark.parametrize("lanes", [1, 2, 4])
@pytest.mark.parametrize("warps", [1, 3, 8, 16])
def test_lanes_per_pass_and_warps(cuda_device, lanes, warps):
    for n_off, N in ((72, 1000), (36, 700), (72, 40)):
        props, scores = synth.make_frames(40, N, n_off, seed=lanes * 17 + warps, groups=4)
        for top_k in (1, 2, 4, 5, 8):
            run_both(props, scores, 50.0, top_k, cuda_device,
                     tuning=dict(variant=3, lanes_per_pass=lanes, stream_warps=warps),
                     ctx=f"lanes={lanes} warps={warps} N={N} No={n_off} top_k={top_k}")


@pytest.mark.parametrize("F", [1, 2, 3, 7, 37, 147, 148, 149, 300, 1000])
def test_frame_counts_around_the_sm_count(cuda_device, F):
    """Fewer frames than SMs: a frame is cut into units of item slots so that every SM has work."""
    for n_off, N in ((72, 1000), (36, 240), (72, 4096)):
        if F * N > 600_000:
            continue
        props, scores = synth.make_frames(F, N, n_off, seed=F + N, groups=3)
        run_both(props, scores, 50.0, 4, cuda_device, tuning=STREAM, ctx=f"F={F} N={N} No={n_off}")


@pytest.mark.parametrize("N", [1, 2, 5, 31, 32, 33, 63, 64, 65, 96, 97, 1023, 1024, 1025, 3000, 8192])
def test_shapes(cuda_device, N):
    for n_off in (72, 36):
        props, scores = synth.make_frames(5, N, n_off, seed=N * 3 + n_off, groups=min(8, max(1, N // 8)))
        for top_k in (1, 4, 8):
            run_both(props, scores, 50.0, top_k, cuda_device, tuning=STREAM, ctx=f"N={N} No={n_off} top_k={top_k}")


@pytest.mark.parametrize("N", [1, 2, 5, 7, 8, 9])
def test_many_frames_of_fewer_proposals_than_one_draw_batch(cuda_device, N):
    """Forced onto the streaming path, the draw cap of a tiny frame (the whole frame for N <= 256) is below one batch of 8 draws
    (found by scripts/soak_random.py: the launch was refused).  (The automatic plan of such a batch is the one-launch kernel.)"""
    props, scores = synth.make_frames(3000, N, 72, seed=N, groups=2)
    assert plan(3000, N, 72, top_k=4)["variant"] == _capi.FUSED_SMALL
    for tuning in (None, STREAM):
        run_both(props, scores, 50.0, 4, cuda_device, tuning=tuning, ctx=f"tiny frames N={N} tuning={tuning}")


@pytest.mark.parametrize("n_off", [72, 36])
def test_edge_frames_ties_and_sort_models(cuda_device, n_off):
    for seed in range(8):
        p, s = synth.edge_frame(n_off, seed=seed)
        for top_k in (1, 4, 8):
            for thr in (50.0, 0.0, -1.0, float("nan"), float("inf")):
                for cap in (0, 8):
                    run_both(p[None], s[None], thr, top_k, cuda_device, tuning=dict(variant=3, select_cap=cap),
                             ctx=f"edge seed={seed} No={n_off} top_k={top_k} thr={thr} cap={cap}")
    for N in (5, 20, 32, 33, 100, 600):
        props, scores = synth.make_frames(5, N, n_off, seed=N, ties=True)
        scores[0, : N // 2] = 1.0
        if N > 4:
            scores[1, 1] = float("nan")
            scores[1, 3] = -float("nan")
            scores[2, ::2] = 0.0
            scores[2, 1::4] = -0.0
        for sm in (0, 1, 2):
            for cap in (0, 8):
                run_both(props, scores, 50.0, 4, cuda_device, sort_model=sm, tuning=dict(variant=3, select_cap=cap),
                         ctx=f"ties N={N} No={n_off} sort_model={sm} cap={cap}")


def test_ragged_and_misaligned(cuda_device):
    F, N = 60, 333          # 333 * 77 words: frames start at every alignment modulo 16 bytes
    g = torch.Generator().manual_seed(3)
    for n_off in (72, 36):
        P = 5 + n_off
        props, scores = synth.make_frames(F, N, n_off, seed=8, groups=3)
        n_valid = torch.randint(0, N + 1, (F,), generator=g, dtype=torch.int32)
        n_valid[0], n_valid[1], n_valid[2], n_valid[3], n_valid[4], n_valid[5] = 0, 1, N, 32, 20, 33
        for cap in (0, 8):
            run_both(props, scores, 50.0, 4, cuda_device, n_valid=n_valid, tuning=dict(variant=3, select_cap=cap),
                     ctx=f"ragged No={n_off} cap={cap}")
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
                got = nms_batched(view, sview, 50.0, 4, n_valid.to(cuda_device), tuning=STREAM)
            st.synchronize()
            assert_same(got, oracle_batched(props, scores, 50.0, 4, n_valid), f"shift={shift} No={n_off}")


def test_subnormal_offsets_and_thresholds(cuda_device):
    for n_off in (72, 36):
        props, scores = synth.make_frames(4, 300, n_off, seed=11)
        props[..., 5:] *= 1e-41
        for thr in (50e-41, 5e-41, 1e-45):
            run_both(props, scores, thr, 4, cuda_device, tuning=STREAM, ctx=f"subnormal No={n_off} thr={thr}")
        props, scores = synth.make_frames(6, 1000, n_off, seed=12, groups=3)
        for thr in (10.0, 20.0, 30.0, 40.0, 50.0):
            run_both(props, scores, thr, 4, cuda_device, tuning=STREAM, ctx=f"No={n_off} thr={thr}")


@pytest.mark.parametrize("N,n_off,top_k,F,groups,tuning", [
    (1000, 72, 4, 4096, 8, None), (1000, 72, 4, 4096, 2, None), (1000, 72, 8, 2048, 3, None), (1000, 36, 8, 4096, 8, None),
    (240, 36, 8, 8192, 4, None), (240, 72, 4, 8192, 2, None), (4096, 72, 4, 512, 4, None),
    (1000, 72, 4, 2048, 2, dict(variant=3, select_cap=8)), (1000, 72, 4, 37, 2, None), (1000, 72, 4, 3, 3, None),
])
def test_repeated_launches_are_identical_and_correct(cuda_device, N, n_off, top_k, F, groups, tuning):
    props, scores = synth.make_frames_chunked(F, N, n_off, seed=N + n_off + top_k, device=cuda_device, groups=groups)
    first = None
    for r in range(6):
        out = nms_batched(props, scores, 50.0, top_k, tuning=tuning)
        if first is None:
            torch.cuda.synchronize()
            first = [t.clone() for t in out]
            idx = torch.arange(0, F, max(1, F // 24))[:24]
            assert_same([t[idx] for t in out], oracle_batched(props[idx].cpu(), scores[idx].cpu(), 50.0, top_k), "soak sample")
        else:
            assert all(torch.equal(a, b) for a, b in zip(out, first)), f"launch {r} differs from launch 0"


@pytest.mark.parametrize("N,n_off", [(33, 72), (64, 36), (64, 72), (100, 36), (130, 72), (240, 36), (240, 72), (300, 36)])
def test_small_frames_many_units_per_cta(cuda_device, N, n_off):
    """Frames with fewer item slots than the CTA has warps: different warps of a CTA work on different frames at the same time
    and drift apart, and with thousands of frames the kept-block ring is reused many times -- the regime in which ring
    requests must go out in ticket order (a hang / corruption here was found and fixed in round 2)."""
    F = 7000
    props, scores = synth.make_frames_chunked(F, N, n_off, seed=N + n_off, device=cuda_device, groups=4)
    ref = nms_batched(props, scores, 50.0, 8, tuning=dict(path=1, variant=_capi.FUSED_REG))
    for tuning in (STREAM, dict(variant=3, stream_warps=16, lanes_per_pass=1), dict(variant=3, stream_warps=5), dict(variant=3, select_cap=8)):
        for rep in range(3):
            got = nms_batched(props, scores, 50.0, 8, tuning=tuning)
            for x, y in zip(got, ref):
                assert torch.equal(x, y), f"N={N} No={n_off} tuning={tuning} rep={rep}"
    idx = torch.arange(0, F, 97)
    assert_same([t[idx] for t in ref], oracle_batched(props[idx].cpu(), scores[idx].cpu(), 50.0, 8), f"oracle sample N={N}")


def test_streaming_and_cluster_kernels_agree_on_a_large_batch(cuda_device):
    F = 2048
    for groups, top_k in ((8, 4), (2, 4), (3, 8)):
        props, scores = synth.make_frames_chunked(F, 1000, 72, seed=groups, device=cuda_device, groups=groups)
        a = nms_batched(props, scores, 50.0, top_k, tuning=STREAM)
        b = nms_batched(props, scores, 50.0, top_k, tuning=dict(path=1, variant=_capi.FUSED_REG))
        c = nms_batched(props, scores, 50.0, top_k, tuning=dict(variant=3, select_cap=8))
        for x, y, z in zip(a, b, c):
            assert torch.equal(x, y) and torch.equal(x, z), f"groups={groups} top_k={top_k}"

"""GPU parity tests: the CUDA op (through the C-ABI) against the CPU oracle, bit-exact on keep / num / parent."""
import numpy as np
import pytest
import torch

from phnet_b200 import synth
from phnet_b200.ops import nms, nms_batched, plan
from tests.util import assert_same, oracle_batched

pytestmark = pytest.mark.gpu


def run_both(props, scores, thr, top_k, dev, n_valid=None, tuning=None, sort_model=0, ctx=""):
    nv = None if n_valid is None else n_valid.to(dev)
    got = nms_batched(props.to(dev), scores.to(dev), thr, top_k, nv, tuning=tuning, sort_model=sort_model)
    torch.cuda.synchronize()
    want = oracle_batched(props, scores, thr, top_k, n_valid, sort_model=sort_model)
    assert_same(got, want, ctx)
    return got


@pytest.mark.parametrize("n_off", [72, 36])
@pytest.mark.parametrize("N", [1, 2, 7, 31, 32, 33, 64, 65, 100, 128, 129, 240, 256, 500, 512, 1000, 1024, 2048])
def test_fused_auto_shapes(cuda_device, N, n_off):
    props, scores = synth.make_frames(6, N, n_off, seed=N * 7 + n_off)
    for top_k in (4, 8):
        run_both(props, scores, 50.0, top_k, cuda_device, ctx=f"N={N} No={n_off} top_k={top_k}")


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("cluster,threads", [(1, 128), (1, 512), (2, 256), (2, 512), (4, 128), (4, 256), (8, 128), (16, 128)])
def test_fused_cluster_sizes(cuda_device, cluster, threads, variant):
    from phnet_b200 import _capi
    for n_off, N in ((72, 700), (36, 900), (72, 240), (36, 1000)):
        tuning = dict(path=1, cluster=cluster, threads=threads, variant=variant)
        try:
            plan(9, N, n_off, tuning)
        except _capi.PhnmsError:
            continue   # this combination does not fit (e.g. 700 rows x 72 offsets in one CTA's registers)
        props, scores = synth.make_frames(9, N, n_off, seed=cluster * 100 + threads + N)
        for top_k in (1, 4, 0):
            run_both(props, scores, 50.0, top_k, cuda_device, tuning=tuning,
                     ctx=f"variant={variant} cluster={cluster} threads={threads} N={N} No={n_off} top_k={top_k}")


def test_both_fused_variants_on_edge_and_ragged(cuda_device):
    F, N = 24, 300
    g = torch.Generator().manual_seed(4)
    for n_off in (72, 36):
        props, scores = synth.make_frames(F, N, n_off, seed=6, ties=True)
        n_valid = torch.randint(0, N + 1, (F,), generator=g, dtype=torch.int32)
        n_valid[0], n_valid[1], n_valid[2], n_valid[3], n_valid[4] = 0, 1, N, 32, 20
        for variant in (1, 2):
            for cluster in (1, 2, 4):
                run_both(props, scores, 50.0, 4, cuda_device, n_valid=n_valid,
                         tuning=dict(path=1, cluster=cluster, variant=variant),
                         ctx=f"ragged variant={variant} cluster={cluster} No={n_off}")
        for seed in range(6):
            p, s = synth.edge_frame(n_off, seed=seed)
            for variant in (1, 2):
                for cluster in (1, 2):
                    for top_k in (0, 4):
                        run_both(p[None], s[None], 50.0, top_k, cuda_device,
                                 tuning=dict(path=1, cluster=cluster, variant=variant),
                                 ctx=f"edge variant={variant} cluster={cluster} No={n_off} seed={seed} top_k={top_k}")


@pytest.mark.parametrize("top_k", [0, 1, 2, 4, 8, 1000, 5000])
@pytest.mark.parametrize("thr", [10.0, 30.0, 50.0])
def test_topk_and_threshold(cuda_device, top_k, thr):
    props, scores = synth.make_frames(4, 1000, 72, seed=int(thr) + top_k)
    run_both(props, scores, thr, top_k, cuda_device, ctx=f"top_k={top_k} thr={thr}")


@pytest.mark.parametrize("n_off", [72, 36, 1, 2, 3, 10, 73, 100, 250])
def test_offset_counts(cuda_device, n_off):
    props, scores = synth.make_frames(3, 200, n_off, seed=n_off)
    run_both(props, scores, 50.0, 4, cuda_device, ctx=f"n_off={n_off}")
    run_both(props, scores, 50.0, 4, cuda_device, tuning=dict(path=2), ctx=f"tiled n_off={n_off}")


@pytest.mark.parametrize("n_off", [72, 36])
def test_edge_frames(cuda_device, n_off):
    for seed in range(12):
        p, s = synth.edge_frame(n_off, seed=seed)
        for top_k in (0, 1, 4, 8, 96):
            for thr in (50.0, 0.0, -1.0, float("nan"), float("inf")):
                for tuning in (None, dict(path=1, cluster=2, threads=128), dict(path=1, variant=1), dict(path=2)):
                    run_both(p[None], s[None], thr, top_k, cuda_device, tuning=tuning,
                             ctx=f"edge seed={seed} No={n_off} top_k={top_k} thr={thr} tuning={tuning}")


def test_subnormal_offsets(cuda_device):
    """fp32 subnormals must survive on every path (no flush-to-zero): the reference's FADDs keep them, and so must the
    packed sub.f32x2 of the register-resident kernel.  A flushed difference would read as distance 0 and suppress."""
    for n_off in (72, 36):
        props, scores = synth.make_frames(4, 300, n_off, seed=11)
        props[..., 5:] *= 1e-41          # |dx| <= 767e-41 = 7.7e-39 < FLT_MIN: every difference and sum is subnormal
        for thr in (50e-41, 5e-41, 1e-45):
            for tuning in (None, dict(path=1, variant=1), dict(path=2)):
                run_both(props, scores, thr, 4, cuda_device, tuning=tuning, ctx=f"subnormal No={n_off} thr={thr} tuning={tuning}")


def test_score_ties_all_sort_models(cuda_device):
    for N in (5, 20, 32, 33, 100, 128, 129, 600):
        props, scores = synth.make_frames(5, N, 72, seed=N, ties=True)
        scores[0, : N // 2] = 1.0
        if N > 4:
            scores[1, 1] = float("nan")
            scores[1, 3] = -float("nan")
            scores[2, ::2] = 0.0
            scores[2, 1::4] = -0.0
        for sm in (0, 1, 2):
            for tuning in (None, dict(path=1, variant=1), dict(path=2)):
                run_both(props, scores, 50.0, 4, cuda_device, sort_model=sm, tuning=tuning,
                         ctx=f"ties N={N} sort_model={sm} tuning={tuning}")


def test_n_valid_ragged(cuda_device):
    F, N = 40, 300
    props, scores = synth.make_frames(F, N, 72, seed=5)
    g = torch.Generator().manual_seed(1)
    n_valid = torch.randint(0, N + 1, (F,), generator=g, dtype=torch.int32)
    n_valid[0], n_valid[1], n_valid[2], n_valid[3] = 0, 1, N, 32
    for tuning in (None, dict(path=1, cluster=4, threads=128), dict(path=2)):
        run_both(props, scores, 50.0, 4, cuda_device, n_valid=n_valid, tuning=tuning, ctx=f"ragged tuning={tuning}")


@pytest.mark.parametrize("N", [64, 257, 1000, 2048])
def test_tiled_path(cuda_device, N):
    props, scores = synth.make_frames(3, N, 72, seed=N + 1)
    for top_k in (4, 0):
        run_both(props, scores, 50.0, top_k, cuda_device, tuning=dict(path=2), ctx=f"tiled N={N} top_k={top_k}")


@pytest.mark.parametrize("N", [2048, 4096, 8192])
def test_stress_sweep_large_n(cuda_device, N):
    # BASELINE config 4: 256..8192 proposals x 72 offsets, thresholds 10..50
    props, scores = synth.make_frames(2, N, 72, seed=N)
    for thr in (10.0, 50.0):
        run_both(props, scores, thr, 4, cuda_device, ctx=f"N={N} thr={thr}")
    run_both(props, scores, 30.0, N, cuda_device, ctx=f"N={N} top_k=N")


def test_too_large_for_a_cluster_falls_back_to_tiled(cuda_device):
    N = 12000
    assert plan(1, N, 72)["path"] == 2
    props, scores = synth.make_frames(1, N, 72, seed=3)
    run_both(props, scores, 50.0, 4, cuda_device, ctx="N=12000")


def test_single_frame_dropin_signature(cuda_device):
    # the exact call get_lanes makes (libs/models/Router4OL.py:460-465)
    props, scores = synth.make_frames(1, 240, 72, seed=11)
    p, s = props[0].to(cuda_device), scores[0].to(cuda_device)
    out = nms(p, s, overlap=50, top_k=4)
    assert isinstance(out, list) and len(out) == 3
    keep, num_to_keep, parent = out
    assert keep.dtype == num_to_keep.dtype == parent.dtype == torch.int64
    assert keep.shape == (240,) and parent.shape == (240,) and num_to_keep.dim() == 0
    assert keep.device == p.device
    kept = keep[:num_to_keep]
    assert kept.numel() == int(num_to_keep) <= 4
    assert (keep[int(num_to_keep):] == 0).all()
    _ = p[kept]
    want = oracle_batched(props, scores, 50.0, 4)
    assert_same(out, want, "drop-in")


def test_non_default_stream_and_misaligned_views(cuda_device):
    props, scores = synth.make_frames(3, 333, 72, seed=2)   # 333 * 77 words: frames are only 4-byte aligned
    big = torch.zeros(3 * 333 * 77 + 3, device=cuda_device)
    for shift in (0, 1, 2, 3):
        view = big[shift: shift + 3 * 333 * 77].view(3, 333, 77)
        view.copy_(props)
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            got = nms_batched(view, scores.to(cuda_device), 50.0, 4)
        st.synchronize()
        assert_same(got, oracle_batched(props, scores, 50.0, 4), f"shift={shift}")


def test_full_size_properties(cuda_device):
    """BASELINE config 2 shape (32 frames x 1000 x 72) and a 4096-frame batch: size-independent properties."""
    F = 4096
    props, scores = synth.make_frames_chunked(F, 1000, 72, seed=9, device=cuda_device)
    keep, num, parent = nms_batched(props, scores, 50.0, 4)
    torch.cuda.synchronize()
    # oracle on a strided sample of frames
    idx = torch.arange(0, F, 128)
    want = oracle_batched(props[idx].cpu(), scores[idx].cpu(), 50.0, 4)
    assert_same((keep[idx], num[idx], parent[idx]), want, "sampled frames of the 4096 batch")
    # (1) determinism / idempotence of the launch
    k2, n2, p2 = nms_batched(props, scores, 50.0, 4)
    assert torch.equal(keep, k2) and torch.equal(num, n2) and torch.equal(parent, p2)
    # (2) kept lanes are in descending score order, the first one is the arg-max score
    ar = torch.arange(F, device=cuda_device)
    ks = torch.gather(scores, 1, keep[:, :4])
    valid = torch.arange(4, device=cuda_device)[None, :] < num[:, None]
    assert ((ks[:, :-1] >= ks[:, 1:]) | ~valid[:, 1:]).all()
    assert (keep[:, 0] == scores.argmax(dim=1)).all()
    # (3) every kept lane is its own parent slot; parents are within [0, num]; padding is zero
    for j in range(4):
        sel = num > j
        assert (parent[ar[sel], keep[sel, j]] == j + 1).all()
    assert (parent >= 0).all() and (parent <= num[:, None]).all()
    assert (keep[:, 4:] == 0).all()
    # (4) re-running on the kept lanes only keeps all of them (kept lanes do not suppress each other)
    sub = torch.gather(props, 1, keep[:, :4, None].expand(F, 4, 77)).contiguous()
    subs = torch.gather(scores, 1, keep[:, :4]).contiguous()
    k3, n3, _ = nms_batched(sub, subs, 50.0, 4, num.to(torch.int32))
    assert torch.equal(n3, num)
    assert ((k3[:, :4] == torch.arange(4, device=cuda_device)[None, :]) | ~valid).all()
    # (5) all three device algorithms agree on a 32-frame clip (config 2)
    for tuning in (dict(path=2), dict(path=1, variant=1), dict(path=1, variant=2, cluster=4)):
        kt, nt, pt = nms_batched(props[:32], scores[:32], 50.0, 4, tuning=tuning)
        assert torch.equal(kt, keep[:32]) and torch.equal(nt, num[:32]) and torch.equal(pt, parent[:32]), tuning


def test_error_behaviour(cuda_device):
    props, scores = synth.make_frames(1, 50, 72, seed=0)
    p, s = props[0].to(cuda_device), scores[0].to(cuda_device)
    with pytest.raises(RuntimeError):
        nms(p.cpu(), s, 50, 4)                       # CHECK_CUDA (nms.cpp:40)
    with pytest.raises(RuntimeError):
        nms(p.t().contiguous().t(), s, 50, 4)        # CHECK_CONTIGUOUS (nms.cpp:41)
    with pytest.raises(RuntimeError):
        nms(p.half(), s.half(), 50, 4)               # AT_DISPATCH_FLOATING_TYPES: float and double only (nms_kernel.cu:171)
    kd, nd, pd = nms(p.double(), s.double(), 50, 4)  # double is dispatched (tests/test_f64.py has the parity cases)
    kf, nf, pf = nms(p, s, 50, 4)
    assert int(nd) == int(nf) and torch.equal(kd, kf) and torch.equal(pd, pf)
    with pytest.raises(RuntimeError):
        nms(p[:, :5].contiguous(), s, 50, 4)         # wrong number of offsets (nms_kernel.cu:154)
    with pytest.raises(RuntimeError):
        nms(torch.zeros(64000, 6, device=cuda_device), torch.zeros(64000, device=cuda_device), 50, 4)  # :158
    keep, num, parent = nms(p[:0], s[:0], 50, 4)     # N == 0: empty result, no launch
    assert keep.numel() == 0 and int(num) == 0 and parent.numel() == 0


def test_scores_of_another_dtype_keep_their_own_order(cuda_device):
    """The reference sorts the scores in whatever dtype they come (nms.cpp:51).  Scores that differ only beyond fp32 precision
    must still be ordered by their double values -- not merged into ties by a cast."""
    import numpy as np
    props, scores = synth.make_frames(1, 300, 72, seed=3, ties=True)
    s64 = scores[0].double()
    bump = torch.linspace(0, 1e-10, 300, dtype=torch.float64)      # later proposals win among fp32-equal (non-zero) scores
    s64 = s64 * (1.0 + bump)
    assert (s64.float() == scores[0]).all()                           # invisible in fp32
    order = np.argsort(-s64.numpy(), kind="stable")
    stand_in = np.empty(300, dtype=np.float32)
    stand_in[order] = -np.arange(300, dtype=np.float32)
    want = oracle_batched(props, torch.from_numpy(stand_in)[None], 50.0, 4, sort_model=2)
    got = nms(props[0].to(cuda_device), s64.to(cuda_device), overlap=50, top_k=4)
    assert_same(got, want, "float64 scores")
    plain = nms(props[0].to(cuda_device), scores[0].to(cuda_device), overlap=50, top_k=4)
    assert not torch.equal(plain[0], got[0]) or True                   # (usually differs: ties fall the other way)


def test_shim_inside_a_reference_style_package_matches_the_oracle(cuda_device, tmp_path):
    """INTEGRATION.md option 0: the reference's `libs/ops/nms.py` left as it is, this repo's `nms_impl.so` in place of the
    reference's compiled module."""
    import importlib
    import sys
    from tests.test_capi_load import _reference_style_package
    root = _reference_style_package(tmp_path)
    sys.path.insert(0, root)
    try:
        ops = importlib.import_module("refstyle_libs.ops")
        for N, n_off, top_k in ((240, 72, 4), (1000, 72, 4), (240, 36, 8), (33, 36, 2)):
            props, scores = synth.make_frames(3, N, n_off, seed=N + top_k, groups=3)
            for f in range(3):
                keep, num, parent = ops.nms(props[f].to(cuda_device), scores[f].to(cuda_device), overlap=50, top_k=top_k)
                assert keep.dtype == torch.int64 and num.dim() == 0 and parent.shape == (N,)
                assert_same((keep, num, parent), oracle_batched(props[f:f + 1], scores[f:f + 1], 50.0, top_k), f"shim N={N}")
    finally:
        sys.path.remove(root)
        for k in [k for k in sys.modules if k.startswith("refstyle_libs")]:
            del sys.modules[k]

"""GPU: device-side `predictions_to_pred` (phnms_decode_lanes_f32) against the reference fixtures and the CPU oracle,
bit-exact on the float64 points; and the clip pipeline get_lanes -> decode_lanes against the oracle chain."""
import numpy as np
import pytest
import torch

from oracle import decode_oracle
from phnet_b200.ops import decode_lanes, get_lanes
from tests.test_decode_oracle import cases

pytestmark = pytest.mark.gpu


def check(points, npoints, meta, want, ctx):
    points, npoints, meta = points.cpu().numpy(), npoints.cpu().numpy(), meta.cpu().numpy()
    for i, w in enumerate(want):
        if w is None:
            assert npoints[i] == 0, f"{ctx} lane {i}: should be skipped"
            continue
        assert npoints[i] == len(w[0]), f"{ctx} lane {i}: {npoints[i]} points, want {len(w[0])}"
        assert np.array_equal(points[i, :npoints[i]], w[0], equal_nan=True), f"{ctx} lane {i}: points differ"
        assert np.array_equal(meta[i], np.float32(w[1])), f"{ctx} lane {i}: metadata differs"


def test_decode_matches_reference_fixtures(cuda_device):
    for key, hdr, inp, pts, cnt, meta in cases():
        L = len(inp)
        rows = torch.from_numpy(inp).to(cuda_device).reshape(1, L, -1).contiguous()
        num = torch.tensor([L], dtype=torch.int64, device=cuda_device)
        p, n, m = decode_lanes(rows, num, 720, 120 if hdr == 7 else 0)
        want = [None if cnt[i] == 0 else (pts[i, :cnt[i]], meta[i]) for i in range(L)]
        check(p[0], n[0], m[0], want, key)


@pytest.mark.parametrize("hdr,n_off", [(6, 72), (6, 36), (7, 36), (7, 72)])
def test_clip_pipeline_get_lanes_then_decode(cuda_device, hdr, n_off):
    T, A, K = 6, 240, 4 if hdr == 6 else 8
    g = torch.Generator().manual_seed(hdr * 100 + n_off)
    out = torch.zeros((T, A, hdr + n_off), dtype=torch.float32)
    out[..., :2] = torch.randn(T, A, 2, generator=g) * 2
    out[..., 2] = torch.rand(T, A, generator=g) * 0.4
    out[..., 3] = torch.rand(T, A, generator=g)
    out[..., 4] = torch.rand(T, A, generator=g)
    out[..., 5] = torch.rand(T, A, generator=g) * 0.9
    if hdr == 7:
        out[..., 6] = torch.rand(T, A, generator=g) * 0.05
    base = torch.rand(T, 6, 1, generator=g) * 0.8 + 0.1
    grp = torch.randint(0, 6, (T, A), generator=g)
    k = torch.arange(n_off, dtype=torch.float32)
    out[..., hdr:] = torch.gather(base.expand(T, 6, n_off), 1, grp[..., None].expand(T, A, n_off)) + 0.002 * k + \
        torch.randn(T, A, n_off, generator=g) * 0.004
    lanes, num, index, mask = get_lanes(out.to(cuda_device), 0.4, 50.0, K)
    points, npoints, meta = decode_lanes(lanes, num, 720, 100)
    torch.cuda.synchronize()
    lanes_h, num_h = lanes.cpu().numpy(), num.cpu().numpy()
    for t in range(T):
        want = decode_oracle.predictions_to_pred(lanes_h[t, :num_h[t]], hdr, 720, 100)
        check(points[t, :num_h[t]], npoints[t, :num_h[t]], meta[t, :num_h[t]], want, f"hdr={hdr} n_off={n_off} frame {t}")
        assert int(npoints[t, num_h[t]:].sum()) == 0
    assert int(num.sum()) > T            # the synthetic clip really keeps lanes

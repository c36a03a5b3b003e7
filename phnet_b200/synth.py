"""Synthetic lane-proposal frames for parity tests and bench.py.

Rows follow what `get_lanes` hands to the op (libs/models/Router4OL.py:454-458):
    [logit0, logit1, start_y (normalised), start_x (px), length (strips), x_0 .. x_{n_off-1} (px)]
Lanes are CLUSTERED, not i.i.d. noise: per frame `groups` base polylines x_k = x0 + s*k + c*k^2 clipped to
[0, img_w-1]; each proposal = one base + N(0, sigma^2) jitter per offset with sigma in {2, 10, 30} px, plus
start_y / length jitter of +-2 strips; 10 % of the proposals are uniform outliers.  (Uniform-random x would give a
mean |dx| of ~256 px >> the 50 px threshold: nothing is ever suppressed and the benchmark degenerates.)
Scores are a random permutation of (1..N)/(N+1) (tie free) or, with ties=True, quantised to 16 levels with some
values saturated at 1.0.
"""
from __future__ import annotations

import torch

IMG_W = 768


def make_frames(F: int, N: int, n_off: int, seed: int = 0, device="cpu", ties: bool = False, groups: int = 8,
                outlier_frac: float = 0.1):
    """Returns (props[F, N, 5+n_off] fp32, scores[F, N] fp32) on `device`."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))

    def rand(*shape):
        return torch.rand(*shape, generator=g, device=dev, dtype=torch.float32)

    def randn(*shape):
        return torch.randn(*shape, generator=g, device=dev, dtype=torch.float32)

    wmax = float(IMG_W - 1)
    scale = 72.0 / n_off
    k = torch.arange(n_off, device=dev, dtype=torch.float32)
    x0 = rand(F, groups) * (wmax - 200.0) + 100.0
    s = (rand(F, groups) * 8.0 - 4.0) * scale
    c = (rand(F, groups) * 0.06 - 0.03) * scale * scale
    base = (x0[..., None] + s[..., None] * k + c[..., None] * k * k).clamp_(0.0, wmax)      # [F, G, No]
    sy_g = rand(F, groups) * 0.35
    len_g = (0.3 + 0.7 * rand(F, groups)) * n_off

    grp = torch.randint(0, groups, (F, N), generator=g, device=dev)
    sig = torch.tensor([2.0, 10.0, 30.0], device=dev)[torch.randint(0, 3, (F, N), generator=g, device=dev)]
    x = torch.gather(base, 1, grp[..., None].expand(F, N, n_off)) + randn(F, N, n_off) * sig[..., None]
    start_y = (torch.gather(sy_g, 1, grp) + (rand(F, N) * 4.0 - 2.0) / max(n_off - 1, 1)).clamp_(0.0, 0.4)
    length = (torch.gather(len_g, 1, grp) + rand(F, N) * 4.0 - 2.0).clamp_(1.0, float(n_off))

    out = rand(F, N) < outlier_frac
    x = torch.where(out[..., None], rand(F, N, n_off) * wmax, x)
    start_y = torch.where(out, rand(F, N) * 0.35, start_y)
    length = torch.where(out, (0.3 + 0.7 * rand(F, N)) * n_off, length)

    props = torch.empty((F, N, 5 + n_off), device=dev, dtype=torch.float32)
    props[..., 0:2] = randn(F, N, 2)
    props[..., 2] = start_y
    props[..., 3] = rand(F, N) * wmax
    props[..., 4] = length
    props[..., 5:] = x

    perm = torch.argsort(rand(F, N), dim=1)
    scores = (perm.to(torch.float32) + 1.0) / float(N + 1)
    if ties:
        scores = torch.floor(scores * 16.0) / 16.0
        scores = torch.where(rand(F, N) < 0.05, torch.ones_like(scores), scores)
    return props, scores


def make_frames_chunked(F: int, N: int, n_off: int, seed: int = 0, device="cuda", chunk: int = 1024, **kw):
    """Same distribution, generated `chunk` frames at a time into one preallocated tensor (bounded temporaries)."""
    dev = torch.device(device)
    props = torch.empty((F, N, 5 + n_off), device=dev, dtype=torch.float32)
    scores = torch.empty((F, N), device=dev, dtype=torch.float32)
    for i, f0 in enumerate(range(0, F, chunk)):
        f1 = min(F, f0 + chunk)
        p, s = make_frames(f1 - f0, N, n_off, seed=seed * 100003 + i, device=dev, **kw)
        props[f0:f1] = p
        scores[f0:f1] = s
    return props, scores


def make_head_output(T: int, A: int, n_off: int, hdr: int, seed: int, groups: int = 8, logit_grid: float = 0.0):
    """Raw detector-head output as `get_lanes` receives it (libs/models/Router4OL.py:437-447): [T, A, hdr + n_off] rows
    (logit0, logit1, start_y, start_x, theta, length, [invalid_length when hdr == 7], x_0 ..), positions normalised to [0, 1].
    logit_grid > 0 rounds the logits to that grid: the softmax scores are then well separated from each other and from any
    threshold, so results do not depend on the last bit of a softmax implementation (CPU-generated fixtures)."""
    props, _ = make_frames(T, A, n_off, seed=seed, groups=groups)
    g = torch.Generator().manual_seed(seed + 1)
    out = torch.empty((T, A, hdr + n_off), dtype=torch.float32)
    logits = torch.randn((T, A, 2), generator=g) * 2.0
    out[..., 0:2] = torch.round(logits / logit_grid) * logit_grid if logit_grid > 0 else logits
    out[..., 2] = props[..., 2]
    out[..., 3] = props[..., 3] / 767.0
    out[..., 4] = torch.rand((T, A), generator=g)
    out[..., 5] = props[..., 4] / (n_off - 1)
    if hdr == 7:
        out[..., 6] = torch.rand((T, A), generator=g) * 0.2
    out[..., hdr:] = props[..., 5:] / 767.0
    return out


def edge_frame(n_off: int, seed: int = 0):
    """One small frame full of the awkward cases the reference code path contains (SURVEY.md section 8d edge suite):
    negative start_y (the `unsigned char` counter wrap), length in {-3, 0, 0.4, 1, 200}, NaN / Inf in x and in the
    header, identical rows, score ties.  Returns (props[N, 5+n_off], scores[N]) on the CPU."""
    g = torch.Generator().manual_seed(seed)
    base, sc = make_frames(1, 96, n_off, seed=seed + 17, ties=True)
    p, sc = base[0].clone(), sc[0].clone()
    n = p.shape[0]
    r = lambda lo, hi: int(torch.randint(lo, hi, (1,), generator=g))  # noqa: E731
    for v in (-0.1, -0.05, -0.001, -1.0, -3.4, -7.0):          # start in [-5,-1] pulls the header into the sum;
        p[r(0, n), 2] = v                                        # start <= -6 wraps the u8 counter past the row
    for v in (-3.0, 0.0, 0.4, 1.0, 200.0, float("nan"), float("inf")):
        p[r(0, n), 4] = v
    for v in (float("nan"), float("inf"), -float("inf"), 1e30):
        p[r(0, n), 5 + r(0, n_off)] = v
    p[r(0, n), 2] = float("nan")
    p[r(0, n), 2] = 1e12
    p[r(0, n), 2] = -1e12
    for _ in range(6):                                         # identical rows (distance 0)
        p[r(0, n)] = p[r(0, n)]
    p[3] = p[2]
    sc[3] = sc[2]
    return p.contiguous(), sc.contiguous()

"""Generates tests/golden/decode_ref.npz by running the REFERENCE's own `predictions_to_pred` bodies on seeded inputs.

The reference modules cannot be imported here (mmcv etc. are missing), so the two method bodies are cut out of the source
files where they lie under /root/reference with `ast` and executed against a stub `self` (prior_ys, n_strips) and a stub
`Lane` that records what the reference passes to its constructor.  Nothing is copied into this repository.

    python tests/golden/make_decode_fixtures.py        (needs /root/reference; CPU only)
"""
import ast
import os
import textwrap

import numpy as np
import torch

REF = "/root/reference/libs/models"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "decode_ref.npz")
if not hasattr(np, "bool"):
    np.bool = bool          # the reference was written against numpy < 1.24 (Router4OLV2.py:385)


def method_source(path, name):
    src = open(path).read()
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.FunctionDef) and node.name == name:
            return textwrap.dedent(ast.get_source_segment(src, node))
    raise KeyError(name)


class RecordingLane:
    def __init__(self, points=None, metadata=None, **kw):
        self.points = np.array(points, dtype=np.float64)
        self.metadata = {k: float(v) for k, v in metadata.items()}


class StubSelf:
    def __init__(self, n_off):
        self.n_offsets = n_off
        self.n_strips = n_off - 1
        self.prior_ys = torch.linspace(1, 0, steps=n_off, dtype=torch.float32)      # Router4OLV2.py:61


def make_inputs(hdr, n_off, L, seed):
    g = torch.Generator().manual_seed(seed)
    p = torch.zeros((L, hdr + n_off), dtype=torch.float32)
    p[:, :2] = torch.randn(L, 2, generator=g)
    p[:, 2] = torch.rand(L, generator=g) * 0.6 - 0.05                 # start_y, a few below 0
    p[:, 3] = torch.rand(L, generator=g)
    p[:, 4] = torch.rand(L, generator=g)
    p[:, 5] = torch.round(torch.rand(L, generator=g) * (n_off + 6) - 3)   # rounded length, some <= 0, some > n_off
    if hdr == 7:
        p[:, 6] = torch.round(torch.rand(L, generator=g) * 8 - 2)         # rounded invalid length, some negative
    base = torch.rand(L, 1, generator=g) * 0.8 + 0.1
    slope = (torch.rand(L, 1, generator=g) - 0.5) * 0.03
    k = torch.arange(n_off, dtype=torch.float32)[None]
    p[:, hdr:] = base + slope * k + torch.randn(L, n_off, generator=g) * 0.01
    # x leaving the image at either end, exact 0 / 1, a NaN and an Inf
    p[1, hdr:hdr + 5] = -0.01
    p[2, hdr + 3] = 1.5
    p[3, hdr:hdr + 2] = torch.tensor([0.0, 1.0])
    p[4, hdr + n_off // 2] = float("nan")
    p[5, hdr + 1] = float("inf")
    p[6, 5] = 0.0
    p[7, 5] = 1.0
    p[8, 5] = -2.0
    p[9, 2] = 0.5 / (n_off - 1)      # round-half-even cases of the start
    p[10, 2] = 1.5 / (n_off - 1)
    p[11, 2] = 2.5 / (n_off - 1)
    return p


def main():
    fns = {6: method_source(os.path.join(REF, "Router4OLV2.py"), "predictions_to_pred"),
           7: method_source(os.path.join(REF, "RouterV4.py"), "predictions_to_pred")}
    out = {}
    for hdr in (6, 7):
        ns = {"torch": torch, "np": np, "Lane": RecordingLane}
        exec(fns[hdr], ns)
        for n_off in (72, 36):
            for seed in range(3):
                L = 40
                p = make_inputs(hdr, n_off, L, seed * 10 + hdr + n_off)
                # the reference skips lanes silently, so it is run one lane at a time to keep the alignment
                pts = np.zeros((L, n_off, 2), dtype=np.float64)
                cnt = np.zeros((L,), dtype=np.int32)
                meta = np.zeros((L, 3), dtype=np.float32)
                for i in range(L):
                    lanes = ns["predictions_to_pred"](StubSelf(n_off), p[i:i + 1].clone(), 720, 120 if hdr == 7 else 0)
                    if lanes:
                        q = lanes[0].points
                        cnt[i] = len(q)
                        pts[i, :len(q)] = q
                        m = lanes[0].metadata
                        meta[i] = (m["start_x"], m["start_y"], m["conf"])
                key = f"h{hdr}_n{n_off}_s{seed}"
                out[key + "_in"] = p.numpy()
                out[key + "_pts"] = pts
                out[key + "_cnt"] = cnt
                out[key + "_meta"] = meta
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, len(out) // 4, "cases")


if __name__ == "__main__":
    main()

"""Golden fixtures for DOUBLE boxes.  RUN ON THE GPU BOX (needs a B200 and oracle/_ref/*.so):

    python tests/golden/make_ref_fixtures_f64.py gpurun_out/golden

Runs the reference's own CUDA op (its nms_kernel<double> instantiation, libs/ops/csrc/nms_kernel.cu:171, compiled from
/root/reference by oracle/build_ref.py) on seeded float64 inputs and stores inputs, torch's CUDA ordering of the float64
scores and the op's outputs in ref_nms_b200_f64.npz (also written next to this script so that the tests of the same run
see it).  Includes start_y values that sit on the rounding boundary of `a[2] * N_STRIPS + 0.5`, where the fused
multiply-add the reference compiles to and a separate multiply + add disagree.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_op  # noqa: E402
from phnet_b200 import synth  # noqa: E402


def cases():
    out = []
    for n_off in (72, 36):
        for (N, seed, ties) in ((300, 0, False), (240, 1, True), (100, 2, True), (33, 3, True), (20, 4, True), (7, 5, False),
                                (129, 6, True)):
            p, s = synth.make_frames(1, N, n_off, seed=seed * 17 + n_off, ties=ties)
            g = torch.Generator().manual_seed(seed)
            pd = p[0].double() + torch.rand(p[0].shape, generator=g, dtype=torch.float64) * 1e-9      # genuinely double
            sd = s[0].double() + (0 if ties else 1) * torch.rand(s[0].shape, generator=g, dtype=torch.float64) * 1e-12
            # start_y on the rounding boundary of y * n_strips + 0.5 (integer boundaries k: y = (k - 0.5) / n_strips +- ulps)
            ns = float(n_off - 1)
            for r in range(min(N, 24)):
                k = 1 + r % 20
                y = np.float64((k - 0.5) / ns)
                for _ in range(r % 5):
                    y = np.nextafter(y, np.float64(1.0) if r % 2 else np.float64(0.0))
                pd[r, 2] = float(y)
            out.append((f"f64_N{N}_No{n_off}_s{seed}", pd.contiguous(), sd.contiguous()))
        for seed in range(3):
            p, s = synth.edge_frame(n_off, seed=seed)
            out.append((f"f64_edge_No{n_off}_s{seed}", p.double().contiguous(), s.double().contiguous()))
    return out


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    dev = torch.device("cuda:0")
    blob = {}
    for name, p, s in cases():
        pc, sc = p.to(dev), s.to(dev)
        blob[name + "/props"] = p.numpy()
        blob[name + "/scores"] = s.numpy()
        blob[name + "/order"] = torch.sort(sc, 0, True)[1].cpu().numpy()
        for thr, top_k in ((50.0, 4), (20.0, 0), (35.0, p.shape[0])):
            keep, num, parent = ref_op.nms(pc, sc, thr, top_k)
            torch.cuda.synchronize()
            tag = f"{name}/thr{thr:g}_k{top_k}"
            blob[tag + "/keep"] = keep.cpu().numpy()
            blob[tag + "/num"] = num.cpu().numpy()
            blob[tag + "/parent"] = parent.cpu().numpy()
    for d in (out_dir, os.path.dirname(os.path.abspath(__file__))):
        np.savez_compressed(os.path.join(d, "ref_nms_b200_f64.npz"), **blob)
    print("wrote ref_nms_b200_f64.npz:", len(cases()), "frames")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")

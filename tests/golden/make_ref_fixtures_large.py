"""Golden fixtures at the large end of the stress sweep (BASELINE config 4).  RUN ON THE GPU BOX (needs a B200 and oracle/_ref):

    python tests/golden/make_ref_fixtures_large.py gpurun_out/golden

Same recipe as make_ref_fixtures.py -- the reference's own CUDA op (oracle/_ref, built from /root/reference/libs/ops/csrc) and
torch's CUDA sort on seeded inputs -- for N = 2048 (72 offsets), N = 4096 and N = 8192 (36 offsets, to keep the file small),
thresholds 10 .. 50, top_k in {0, 4, 8, N}, plus ragged prefixes of the same frames (the n_valid case of the batched op).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_op  # noqa: E402
from phnet_b200 import synth  # noqa: E402


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    dev = torch.device("cuda:0")
    blob = {}
    for (N, n_off, seed, ties, groups) in ((2048, 72, 0, False, 3), (4096, 36, 1, True, 8), (8192, 36, 2, False, 2)):
        p, s = synth.make_frames(1, N, n_off, seed=seed * 7 + N, ties=ties, groups=groups)
        p, s = p[0], s[0]
        name = f"synth_N{N}_No{n_off}_s{seed}"
        blob[name + "/props"] = p.numpy()
        blob[name + "/scores"] = s.numpy()
        for nv in (N, N - 333, 1500):       # the frame and two prefixes of it
            pc, sc = p[:nv].to(dev).contiguous(), s[:nv].to(dev).contiguous()
            blob[f"{name}/n{nv}/order"] = torch.sort(sc, 0, True)[1].cpu().numpy()
            for thr, top_k in ((50.0, 4), (10.0, 8), (20.0, 0), (30.0, 1), (40.0, nv)):
                keep, num, parent = ref_op.nms(pc, sc, thr, top_k)
                torch.cuda.synchronize()
                tag = f"{name}/n{nv}/thr{thr:g}_k{top_k}"
                blob[tag + "/keep"] = keep.cpu().numpy().astype(np.int32)      # (indices fit; halves the file)
                blob[tag + "/num"] = num.cpu().numpy()
                blob[tag + "/parent"] = parent.cpu().numpy().astype(np.int32)
    np.savez_compressed(os.path.join(out_dir, "ref_nms_b200_large.npz"), **blob)
    print("wrote", os.path.join(out_dir, "ref_nms_b200_large.npz"))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")

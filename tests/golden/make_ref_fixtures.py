"""Generates the committed golden fixtures.  RUN ON THE GPU BOX (needs a B200 and oracle/_ref/*.so):

    python tests/golden/make_ref_fixtures.py gpurun_out/golden

It runs (a) the reference's own CUDA op, compiled unmodified-but-for-two-tokens from /root/reference/libs/ops/csrc
by oracle/build_ref.py, and (b) torch's CUDA `scores.sort(0, True)` (the un-vendored ordering the reference
delegates to, libs/ops/csrc/nms.cpp:51) on seeded inputs, and stores inputs + outputs as .npz.  The files are then
copied into tests/golden/ and committed; the CPU oracle and the CUDA op are both tested against them.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_op  # noqa: E402
from phnet_b200 import synth  # noqa: E402


def ref_cases():
    cases = []
    for n_off in (72, 36):
        for (N, seed, ties) in ((1000, 0, False), (1000, 1, True), (240, 2, False), (100, 3, True), (33, 4, True),
                                (32, 5, True), (20, 6, True), (7, 7, False), (2, 8, True), (129, 9, True), (600, 10, False)):
            p, s = synth.make_frames(1, N, n_off, seed=seed * 31 + n_off, ties=ties)
            cases.append((f"synth_N{N}_No{n_off}_s{seed}", p[0], s[0]))
        for seed in range(4):
            p, s = synth.edge_frame(n_off, seed=seed)
            cases.append((f"edge_No{n_off}_s{seed}", p, s))
    return cases


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    dev = torch.device("cuda:0")
    blob = {}
    names = []
    for name, p, s in ref_cases():
        n_off = p.shape[1] - 5
        pc, sc = p.to(dev).contiguous(), s.to(dev).contiguous()
        blob[name + "/props"] = p.numpy()
        blob[name + "/scores"] = s.numpy()
        order = torch.sort(sc, 0, True)[1]
        blob[name + "/order"] = order.cpu().numpy()
        for thr, top_k in ((50.0, 4), (50.0, 8), (20.0, 0), (50.0, 1), (35.0, p.shape[0])):
            keep, num, parent = ref_op.nms(pc, sc, thr, top_k)
            torch.cuda.synchronize()
            tag = f"{name}/thr{thr:g}_k{top_k}"
            blob[tag + "/keep"] = keep.cpu().numpy()
            blob[tag + "/num"] = num.cpu().numpy()
            blob[tag + "/parent"] = parent.cpu().numpy()
        names.append(name)
    np.savez_compressed(os.path.join(out_dir, "ref_nms_b200.npz"), **blob)

    # torch CUDA sort tie / NaN / signed-zero behaviour for every size class
    g = torch.Generator().manual_seed(123)
    sort_blob = {}
    for N in list(range(1, 40)) + [63, 64, 65, 100, 127, 128, 129, 130, 255, 256, 257, 1000, 1024, 2048, 4095, 4096, 4097,
                                   5000, 8192, 20000]:
        for variant in range(4):
            s = torch.floor(torch.rand(N, generator=g) * 8.0) / 8.0
            if variant >= 1 and N > 3:
                s[torch.randint(0, N, (max(1, N // 10),), generator=g)] = 1.0
                s[torch.randint(0, N, (max(1, N // 16),), generator=g)] = 0.0
                s[torch.randint(0, N, (max(1, N // 16),), generator=g)] = -0.0
            if variant >= 2 and N > 3:
                s[torch.randint(0, N, (max(1, N // 12),), generator=g)] = float("nan")
                s[torch.randint(0, N, (max(1, N // 20),), generator=g)] = float("inf")
                s[torch.randint(0, N, (max(1, N // 20),), generator=g)] = -float("inf")
            if variant >= 3 and N > 3:
                neg_nan = torch.tensor([0xFFC00000], dtype=torch.int64).to(torch.int32).view(torch.float32)[0]
                s[torch.randint(0, N, (max(1, N // 12),), generator=g)] = neg_nan
                s[torch.randint(0, N, (max(1, N // 8),), generator=g)] *= -1.0
            order = torch.sort(s.to(dev), 0, True)[1].cpu()
            order_stable = torch.sort(s.to(dev), dim=0, descending=True, stable=True)[1].cpu()
            sort_blob[f"N{N}_v{variant}/scores_bits"] = s.view(torch.int32).numpy()
            sort_blob[f"N{N}_v{variant}/order"] = order.numpy()
            sort_blob[f"N{N}_v{variant}/order_stable"] = order_stable.numpy()
    np.savez_compressed(os.path.join(out_dir, "torch_cuda_sort.npz"), **sort_blob)
    with open(os.path.join(out_dir, "README.txt"), "w") as f:
        f.write(f"generated on {torch.cuda.get_device_name(0)} with torch {torch.__version__}\n")
        f.write("ref_nms_b200.npz: outputs of the reference CUDA op (oracle/_ref) on seeded inputs\n")
        f.write("torch_cuda_sort.npz: torch.sort(descending=True) on CUDA, tie-heavy vectors\n")
    print("wrote", out_dir, len(names), "cases")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")

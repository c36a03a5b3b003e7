"""The CUDA op against the REFERENCE itself, on the GPU box (closes the parity chain there, where the driver runs `-m gpu`):

  1. every case of the committed golden fixtures -- outputs of the reference's own CUDA op (libs/ops/csrc/nms_kernel.cu:26-192,
     nms.cpp:44-57) captured on a B200 by tests/golden/make_ref_fixtures*.py -- through every device algorithm of this repo;
  2. the CPU oracle against the same fixtures (the check tests/test_oracle.py makes on the CPU box, repeated here so that the
     GPU run does not depend on a deselected test);
  3. when oracle/_ref/phnet_ref_nms_{36,72}.so load (they are built by __graft_entry__.build() where /root/reference exists
     and travel with the snapshot): the reference op run LIVE on a few hundred seeded frames, all three outputs compared.

Bit-exact on keep / num_to_keep / parent_object_index.
"""
import os

import numpy as np
import pytest
import torch

from oracle import oracle, ref_op
from phnet_b200 import _capi, synth
from phnet_b200.ops import nms, nms_batched

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

TUNINGS = {
    "auto": None,
    "stream": dict(variant=_capi.FUSED_STREAM),
    "stream_cap8": dict(variant=_capi.FUSED_STREAM, select_cap=8),
    "cluster_reg": dict(path=1, variant=_capi.FUSED_REG),
    "cluster_smem": dict(path=1, variant=_capi.FUSED_SMEM),
    "tiled": dict(path=2),
}


def _same(got, keep, num, parent, ctx):
    k, n, p = got
    assert int(n) == int(num), f"{ctx}: num_to_keep {int(n)} != {int(num)}"
    assert np.array_equal(k.cpu().numpy(), keep.astype(np.int64)), f"{ctx}: keep differs"
    assert np.array_equal(p.cpu().numpy(), parent.astype(np.int64)), f"{ctx}: parent differs"


def _runs(z, prefix):
    for key in sorted(k for k in z.files if k.startswith(prefix + "/thr") and k.endswith("/keep")):
        tag = key[: -len("/keep")]
        thr = float(tag.split("/thr")[1].split("_k")[0])
        top_k = int(tag.split("_k")[1])
        yield tag, thr, top_k


def test_cuda_op_matches_reference_golden(cuda_device):
    z = np.load(os.path.join(GOLD, "ref_nms_b200.npz"))
    checked = 0
    for nm in sorted({k.split("/")[0] for k in z.files}):
        p = torch.from_numpy(z[nm + "/props"]).to(cuda_device)
        s = torch.from_numpy(z[nm + "/scores"]).to(cuda_device)
        for tag, thr, top_k in _runs(z, nm):
            for tname, tune in TUNINGS.items():
                if tname.startswith("stream") and not 1 <= top_k <= 8:
                    continue
                _same(nms(p, s, overlap=thr, top_k=top_k, tuning=tune), z[tag + "/keep"], z[tag + "/num"], z[tag + "/parent"],
                      f"{tag} [{tname}]")
                checked += 1
    assert checked >= 700


def test_cuda_op_matches_reference_golden_large_and_ragged(cuda_device):
    path = os.path.join(GOLD, "ref_nms_b200_large.npz")
    z = np.load(path)
    checked = 0
    for nm in sorted({k.split("/")[0] for k in z.files}):
        props, scores = z[nm + "/props"], z[nm + "/scores"]
        N = props.shape[0]
        p = torch.from_numpy(props).to(cuda_device)
        s = torch.from_numpy(scores).to(cuda_device)
        for nvtag in sorted({k.split("/")[1] for k in z.files if k.startswith(nm + "/n")}):
            nv = int(nvtag[1:])
            assert np.array_equal(oracle.order(scores[:nv]), z[f"{nm}/{nvtag}/order"]), f"{nm}/{nvtag}: oracle order != torch CUDA sort"
            for tag, thr, top_k in _runs(z, f"{nm}/{nvtag}"):
                keep, num, parent = z[tag + "/keep"], z[tag + "/num"], z[tag + "/parent"]
                ok, on, op = oracle.nms(props[:nv], scores[:nv], thr, top_k)
                assert on == int(num) and np.array_equal(ok, keep) and np.array_equal(op, parent), f"{tag}: oracle differs from the reference"
                for tname in ("auto", "stream_cap8", "cluster_reg", "tiled"):
                    if tname.startswith("stream") and not 1 <= top_k <= 8:
                        continue
                    # (a) the prefix as a frame of its own, (b) the whole frame with n_valid = the prefix length
                    _same(nms(p[:nv].contiguous(), s[:nv].contiguous(), overlap=thr, top_k=top_k, tuning=TUNINGS[tname]),
                          keep, num, parent, f"{tag} [{tname}]")
                    k, n, par = nms_batched(p[None], s[None], thr, top_k, torch.tensor([nv], dtype=torch.int32, device=cuda_device),
                                            tuning=TUNINGS[tname])
                    _same((k[0, :nv], n[0], par[0, :nv]), keep, num, parent, f"{tag} [{tname}, n_valid]")
                    assert not k[0, nv:].any() and not par[0, nv:].any(), f"{tag} [{tname}]: padding rows must stay zero"
                    checked += 2
    assert checked >= 150


def test_oracle_matches_reference_golden_on_the_gpu_box():
    """tests/test_oracle.py::test_oracle_matches_reference_cuda_op_golden, under the gpu marker."""
    z = np.load(os.path.join(GOLD, "ref_nms_b200.npz"))
    checked = 0
    for nm in sorted({k.split("/")[0] for k in z.files}):
        p, s = z[nm + "/props"], z[nm + "/scores"]
        assert (oracle.order(s) == z[nm + "/order"]).all(), f"{nm}: order differs from torch CUDA sort"
        for tag, thr, top_k in _runs(z, nm):
            for lazy in (False, True):
                keep, num, parent = oracle.nms(p, s, thr, top_k, lazy=lazy)
                assert num == int(z[tag + "/num"]) and (keep == z[tag + "/keep"]).all() and (parent == z[tag + "/parent"]).all(), (tag, lazy)
                checked += 1
    assert checked >= 300


def _live_frames(n_off):
    """(name, props[N, 5+n_off], scores[N]) : generator default, road-like, tie-heavy, tiny, edge and large frames."""
    for N, seed, ties, groups, outl in ((1000, 0, False, 8, 0.1), (1000, 1, True, 2, 0.1), (1000, 2, False, 3, 0.0),
                                        (240, 3, False, 4, 0.1), (240, 4, True, 1, 0.0), (100, 5, True, 8, 0.1), (33, 6, True, 4, 0.1),
                                        (32, 7, True, 4, 0.1), (20, 8, True, 2, 0.1), (7, 9, False, 2, 0.1), (2, 10, True, 1, 0.1),
                                        (1, 11, False, 1, 0.1), (2048, 12, False, 4, 0.05), (4096, 13, True, 8, 0.1)):
        for rep in range(3 if N <= 1000 else 1):
            p, s = synth.make_frames(1, N, n_off, seed=1000 * seed + rep + n_off, ties=ties, groups=groups, outlier_frac=outl)
            yield f"N{N}_s{seed}_{rep}", p[0], s[0]
    for seed in range(6):
        p, s = synth.edge_frame(n_off, seed=100 + seed)
        yield f"edge_{seed}", p, s


@pytest.mark.parametrize("n_off", [72, 36])
def test_cuda_op_matches_live_reference_op(cuda_device, n_off):
    if ref_op.path(n_off) is None:
        pytest.skip("oracle/_ref is not built here (needs /root/reference at build time)")
    checked = 0
    for name, p, s in _live_frames(n_off):
        pc, sc = p.to(cuda_device).contiguous(), s.to(cuda_device).contiguous()
        N = p.shape[0]
        for thr in (10.0, 20.0, 30.0, 40.0, 50.0):
            for top_k in (0, 1, 4, 8, N):
                if N > 1000 and top_k in (0, N) and thr != 30.0:
                    continue      # (the reference's serial collect is slow with unbounded top_k on large frames)
                rk, rn, rp = ref_op.nms(pc, sc, thr, top_k)
                torch.cuda.synchronize()
                for tname in ("auto", "stream_cap8", "cluster_reg"):
                    if tname.startswith("stream") and not 1 <= top_k <= 8:
                        continue
                    k, n, par = nms(pc, sc, overlap=thr, top_k=top_k, tuning=TUNINGS[tname])
                    ctx = f"{name} No={n_off} thr={thr} top_k={top_k} [{tname}]"
                    assert int(n) == int(rn), f"{ctx}: num_to_keep {int(n)} != reference {int(rn)}"
                    assert torch.equal(k, rk), f"{ctx}: keep differs from the live reference op"
                    assert torch.equal(par, rp), f"{ctx}: parent differs from the live reference op"
                    checked += 1
    assert checked >= 1500

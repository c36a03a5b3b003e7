"""Double precision boxes (the reference's nms_kernel<double>, libs/ops/csrc/nms_kernel.cu:171): the CPU restatement and
the CUDA op against outputs of the reference's own double kernels run on a B200 (tests/golden/ref_nms_b200_f64.npz, made
by tests/golden/make_ref_fixtures_f64.py).  Bit-exact on keep / num / parent."""
import os

import numpy as np
import pytest

from oracle import oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_nms_b200_f64.npz")


def cases():
    z = np.load(GOLD)
    names = sorted({k.split("/")[0] for k in z.files})
    for name in names:
        runs = sorted({k.split("/")[1] for k in z.files if k.startswith(name + "/thr")})
        yield name, z[name + "/props"], z[name + "/scores"], z[name + "/order"], \
            [(r, z[f"{name}/{r}/keep"], int(z[f"{name}/{r}/num"]), z[f"{name}/{r}/parent"]) for r in runs]


def parse(run):
    thr, k = run[3:].split("_k")
    return float(thr), int(k)


def test_oracle_f64_matches_reference_double_kernels():
    n = 0
    for name, props, scores, order, runs in cases():
        for run, keep, num, parent in runs:
            thr, top_k = parse(run)
            k, m, p = oracle.nms_f64(props, order, thr, top_k)
            assert m == num and np.array_equal(k, keep) and np.array_equal(p, parent), f"{name} {run}"
            n += 1
    assert n >= 60


@pytest.mark.gpu
def test_cuda_f64_matches_reference_double_kernels(cuda_device):
    import torch
    from phnet_b200.ops import nms, nms_batched
    for name, props, scores, order, runs in cases():
        p = torch.from_numpy(props).to(cuda_device)
        s = torch.from_numpy(scores).to(cuda_device)
        assert np.array_equal(torch.sort(s, 0, True)[1].cpu().numpy(), order), f"{name}: torch's ordering changed"
        for run, keep, num, parent in runs:
            thr, top_k = parse(run)
            k, m, par = nms(p, s, overlap=thr, top_k=top_k)
            assert k.dtype == torch.int64 and m.dim() == 0
            assert int(m) == num and np.array_equal(k.cpu().numpy(), keep) and np.array_equal(par.cpu().numpy(), parent), \
                f"{name} {run}"
    # batched form: F frames of one shape, tie-free scores
    from phnet_b200 import synth
    pr, sc = synth.make_frames(5, 200, 72, seed=3)
    pd, sd = pr.double().to(cuda_device), sc.double().to(cuda_device)
    kb, nb, pb = nms_batched(pd, sd, 50.0, 4)
    for f in range(5):
        idx = np.argsort(-sc[f].numpy().astype(np.float64), kind="stable")
        k, m, par = oracle.nms_f64(pr[f].numpy().astype(np.float64), idx, 50.0, 4)
        assert int(nb[f]) == m and np.array_equal(kb[f].cpu().numpy(), k) and np.array_equal(pb[f].cpu().numpy(), par)
    with pytest.raises(RuntimeError):
        nms(p.half(), s.half(), overlap=50, top_k=4)


@pytest.mark.gpu
def test_ragged_double_batches_equal_per_frame_calls(cuda_device):
    """n_valid with float64 boxes: frame f of the batch equals the one-frame call on its first n_valid[f] rows (that call is what
    the fixtures above pin against the reference's nms_kernel<double>)."""
    import torch
    from phnet_b200 import synth
    from phnet_b200.ops import nms, nms_batched
    F, N = 7, 150
    for n_off in (72, 36):
        props, scores = synth.make_frames(F, N, n_off, seed=5 + n_off, groups=3)
        p, s = props.double().to(cuda_device), scores.double().to(cuda_device)
        n_valid = torch.tensor([150, 0, 1, 40, 33, 149, 77], dtype=torch.int32, device=cuda_device)
        keep, num, parent = nms_batched(p, s, 50.0, 4, n_valid)
        for f in range(F):
            n = int(n_valid[f])
            k1, n1, p1 = nms(p[f, :n].contiguous(), s[f, :n].contiguous(), 50.0, 4)
            assert int(num[f]) == int(n1), f"frame {f}"
            assert torch.equal(keep[f, :n], k1) and torch.equal(parent[f, :n], p1), f"frame {f}"
            assert bool((keep[f, n:] == 0).all()) and bool((parent[f, n:] == 0).all())

"""GPU tests of the fused collection of the kept lanes (phnms_forward_collect_f32): the compact records the kernels store
must equal pack_kept(keep, num) of the same call, in every destination buffer, on every path."""
import os
import subprocess
import sys

import pytest
import torch

from phnet_b200 import _capi, peer, sharding, synth
from phnet_b200.ops import nms_batched, plan
from tests.util import assert_same, oracle_batched

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("tuning", [None, dict(path=1, variant=1), dict(path=2), dict(path=1, cluster=4, threads=256)])
@pytest.mark.parametrize("top_k", [1, 4, 8, 16, 20])
def test_records_match_keep_and_num(cuda_device, tuning, top_k):
    dev = cuda_device
    for N, n_off, F in ((1000, 72, 37), (240, 36, 50), (20, 72, 9)):
        if tuning is not None:
            try:
                plan(F, N, n_off, tuning)
            except _capi.PhnmsError:
                continue     # this override does not fit the shape
        props, scores = synth.make_frames(F, N, n_off, seed=top_k + N)
        row0, rows = 5, F + 11
        bufs = [torch.full((rows, top_k + 1), -7, dtype=torch.int64, device=dev) for _ in range(3)]
        got = nms_batched(props.to(dev), scores.to(dev), 50.0, top_k, tuning=tuning, collect=peer.local_collect(bufs, row0))
        torch.cuda.synchronize()
        assert_same(got, oracle_batched(props, scores, 50.0, top_k), f"collect N={N} top_k={top_k} tuning={tuning}")
        want = sharding.pack_kept(got[0], got[1], top_k)
        for b in bufs:
            assert torch.equal(b[row0:row0 + F], want), f"records differ: N={N} top_k={top_k} tuning={tuning}"
            assert bool((b[:row0] == -7).all()) and bool((b[row0 + F:] == -7).all()), "stores outside the call's rows"


@pytest.mark.parametrize("top_k", [3, 4, 8])
def test_records_of_frames_redone_by_the_resume_pass(cuda_device, top_k):
    """The select kernel stores a record for every frame; a frame it left open (draw cap 8 here) and the streaming pass found
    unfinished is redone by the cluster kernel, which stores the record again -- the final one must be in every buffer."""
    dev = cuda_device
    F, N = 400, 1000
    props, scores = synth.make_frames(F, N, 72, seed=top_k, groups=2, outlier_frac=0.03)
    bufs = [torch.full((F + 3, top_k + 1), -7, dtype=torch.int64, device=dev) for _ in range(2)]
    got = nms_batched(props.to(dev), scores.to(dev), 50.0, top_k, tuning=dict(variant=3, select_cap=8), collect=peer.local_collect(bufs, 2))
    torch.cuda.synchronize()
    assert_same(got, oracle_batched(props, scores, 50.0, top_k), f"collect + resume top_k={top_k}")
    want = sharding.pack_kept(got[0], got[1], top_k)
    for b in bufs:
        assert torch.equal(b[2:2 + F], want) and bool((b[:2] == -7).all()) and bool((b[2 + F:] == -7).all())


def test_ragged_and_empty_frames(cuda_device):
    dev = cuda_device
    F, N = 24, 300
    props, scores = synth.make_frames(F, N, 72, seed=2, ties=True)
    n_valid = torch.randint(0, N + 1, (F,), generator=torch.Generator().manual_seed(3), dtype=torch.int32)
    n_valid[0], n_valid[1], n_valid[2] = 0, 1, 2
    buf = torch.full((F, 5), -1, dtype=torch.int64, device=dev)
    keep, num, _ = nms_batched(props.to(dev), scores.to(dev), 50.0, 4, n_valid.to(dev), collect=peer.local_collect([buf]))
    torch.cuda.synchronize()
    assert torch.equal(buf, sharding.pack_kept(keep, num, 4))
    assert buf[0].tolist() == [0, 0, 0, 0, 0] and int(buf[1, 4]) == 1


def test_records_and_completion_in_one_launch(cuda_device):
    """phnms_collect with signal / wait epochs: the record kernel's last block releases the epoch flag and waits for it
    (here: one rank, its own flag array -- the protocol of phnet_b200.peer.PeerCollector on a single GPU)."""
    import ctypes
    dev = cuda_device
    F, N, top_k = 300, 1000, 4
    props, scores = synth.make_frames(F, N, 72, seed=21, groups=3)
    p, s = props.to(dev), scores.to(dev)
    buf = torch.full((F, top_k + 1), -1, dtype=torch.int64, device=dev)
    flags = torch.zeros(64, dtype=torch.int64, device=dev)       # [0]: epoch flag, [16]: status, [24]: block counter
    for epoch in (1, 2, 3, 7):
        buf.fill_(-1)
        c = peer.local_collect([buf])
        c.signal_epoch, c.wait_epoch, c.timeout_ns = epoch, epoch, int(2e9)
        c.signal_dst[0] = flags.data_ptr()
        c.wait_src = flags.data_ptr()
        c.status = flags.data_ptr() + 16 * 8
        c.sync_counter = flags.data_ptr() + 24 * 8
        keep, num, _ = nms_batched(p, s, 50.0, top_k, collect=c)
        torch.cuda.synchronize()
        assert int(flags[0]) == epoch and int(flags[16]) == 0 and int(flags[24]) == 0
        assert torch.equal(buf, sharding.pack_kept(keep, num, top_k))
    # a wait for an epoch nobody signals times out and reports the slot
    c = peer.local_collect([buf])
    c.wait_epoch, c.timeout_ns = 99, int(2e7)
    c.wait_src, c.status, c.sync_counter = flags.data_ptr(), flags.data_ptr() + 16 * 8, flags.data_ptr() + 24 * 8
    nms_batched(p, s, 50.0, top_k, collect=c)
    torch.cuda.synchronize()
    assert int(flags[16]) & 0xffffffff == 1
    # descriptor errors: a wrong width / too few rows never reach the device
    bad = peer.local_collect([buf])
    bad.width = top_k
    with pytest.raises(_capi.PhnmsError):
        nms_batched(p, s, 50.0, top_k, collect=bad)
    with pytest.raises(_capi.PhnmsError):
        nms_batched(p, s, 50.0, top_k, collect=peer.local_collect([buf[: F - 1]]))
    with pytest.raises(_capi.PhnmsError):
        nms_batched(p, s, 50.0, top_k, collect=peer.local_collect([buf], row0=1))


def test_two_gpu_peer_collection():
    """Two ranks, each storing its records into both ranks' buffers over peer memory (torchrun, NCCL for the plumbing)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tests", "peer_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0 and "peer collection ok" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]

"""The host-buffer entry point (pinned host tensors in, reference-shaped host results out) against the oracle."""
import pytest
import torch

from phnet_b200 import synth
from phnet_b200.ops import HostLaneNMS, nms_host
from tests.util import assert_same, oracle_batched

pytestmark = pytest.mark.gpu


def test_host_pipeline_chunks_and_reuse(cuda_device):
    props, scores = synth.make_frames(37, 300, 72, seed=8)
    pipe = HostLaneNMS(300, 72, chunk_frames=8, device=cuda_device)          # 5 chunks, the last one ragged
    out = pipe(props.pin_memory(), scores.pin_memory(), 50.0, 4)
    torch.cuda.synchronize()
    assert all(not t.is_cuda for t in out)
    want = oracle_batched(props, scores, 50.0, 4)
    assert_same(out, want, "host pipeline")
    assert pipe.launches == 5 and pipe.h2d_bytes == 37 * 300 * 78 * 4 and pipe.d2h_bytes == 37 * (2 * 300 + 1) * 8
    out2 = pipe(props.pin_memory(), scores.pin_memory(), 30.0, 8, out=out)   # staging buffers are reused
    torch.cuda.synchronize()
    assert_same(out2, oracle_batched(props, scores, 30.0, 8), "host pipeline, second call")


def test_nms_host_single_frame_and_errors(cuda_device):
    props, scores = synth.make_frames(1, 120, 36, seed=1)
    keep, num, parent = nms_host(props[0], scores[0], 50, 4, device=cuda_device)
    assert_same((keep, num, parent), oracle_batched(props, scores, 50.0, 4), "nms_host")
    with pytest.raises(RuntimeError):
        HostLaneNMS(120, 36, device=cuda_device)(props.to(cuda_device), scores.to(cuda_device), 50, 4)

"""CUDA-graph replay of the one-frame drop-in call for a fixed shape.

PHNet calls `nms(boxes[N, 5+n_off], scores[N], overlap, top_k)` once per frame (libs/models/Router4OL.py:460-465); the call is one
7 us kernel launch, so what is left per call is host time (~5 us in the pybind shim, ~6.5 us through `phnet_b200.ops.nms`).  When
the shape is fixed -- e.g. the padded `[240, 5+n_off]` proposal block of a frame -- the launch can be captured once and replayed:
`GraphedNMS` owns static input / output buffers, captures the C-ABI call (`phnms_forward_f32`, optionally with `n_valid` so that
the real number of proposals may change from call to call) and replays it.  Results are written to the same output tensors on
every replay: consume them (or copy them) before the next call.
"""
from __future__ import annotations

import torch

from .. import _capi

__all__ = ["GraphedNMS"]


class GraphedNMS:
    def __init__(self, N: int, n_off: int, overlap, top_k, device="cuda:0", sort_model: int = _capi.SORT_TORCH_CUDA,
                 ragged: bool = False):
        self.N, self.n_off, self.top_k = int(N), int(n_off), int(top_k)
        if self.top_k < 0:
            raise TypeError("top_k must be non-negative (unsigned long in the reference, nms.cpp:48)")
        self.dev = torch.device(device)
        d = self.dev
        self.boxes = torch.zeros((self.N, 5 + self.n_off), dtype=torch.float32, device=d)    # static inputs: fill these ...
        self.scores = torch.zeros((self.N,), dtype=torch.float32, device=d)
        self.n_valid = torch.full((1,), self.N, dtype=torch.int32, device=d) if ragged else None
        out = torch.zeros(2 * self.N + 1, dtype=torch.int64, device=d)                       # ... and read these after replay()
        self.keep, self.parent, self.num = out[: self.N], out[self.N: 2 * self.N], out[2 * self.N]
        L = _capi.lib()
        with torch.cuda.device(d):
            nbytes = int(L.phnms_workspace_bytes(1, self.N, self.n_off, None))
            self._ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=d)

            def launch():
                rc = L.phnms_forward_f32(self.boxes.data_ptr(), self.scores.data_ptr(),
                                         self.n_valid.data_ptr() if self.n_valid is not None else None, 1, self.N, self.n_off,
                                         float(overlap), self.top_k, int(sort_model), self.keep.data_ptr(), self.num.data_ptr(),
                                         self.parent.data_ptr(), self._ws.data_ptr() if nbytes else None, nbytes, None,
                                         torch.cuda.current_stream(d).cuda_stream)
                _capi.check(rc)

            side = torch.cuda.Stream(d)
            side.wait_stream(torch.cuda.current_stream(d))
            with torch.cuda.stream(side):
                launch()                       # warm-up outside the capture: per-device attribute set-up happens here
            torch.cuda.current_stream(d).wait_stream(side)
            torch.cuda.synchronize(d)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                launch()

    def replay(self):
        """Run the op on what `self.boxes` / `self.scores` (/ `self.n_valid`) hold now; returns [keep, num_to_keep, parent]."""
        self.graph.replay()
        return [self.keep, self.num, self.parent]

    def __call__(self, boxes: torch.Tensor, scores: torch.Tensor):
        """Copy a frame of at most N proposals into the static buffers (device-to-device, asynchronous) and replay."""
        n = boxes.shape[0]
        if n == self.N:
            self.boxes.copy_(boxes, non_blocking=True)
            self.scores.copy_(scores, non_blocking=True)
        else:
            if self.n_valid is None or n > self.N:
                raise RuntimeError("a frame of another size needs GraphedNMS(..., ragged=True) and at most N proposals")
            self.boxes[:n].copy_(boxes, non_blocking=True)
            self.scores[:n].copy_(scores, non_blocking=True)
        if self.n_valid is not None:
            self.n_valid.fill_(n)
        return self.replay()

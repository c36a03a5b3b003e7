"""Host-buffer entry point: lane NMS on frames that live in (pinned) host memory.

This is the end-to-end call a PHNet-side user makes when proposals are produced off-device or results are consumed
on the host (the reference's get_lanes ends in `.cpu()` decodes, libs/models/Router4OLV2.py:363-404).  Frames are
streamed through the GPU in chunks: H2D copy, the CUDA op (C-ABI `phnms_forward_f32`) and the D2H copy of the
reference-shaped results run on three streams over double-buffered device staging, so PCIe and the kernel overlap.
"""
from __future__ import annotations

import torch

from .nms import nms_batched


class HostLaneNMS:
    """Reusable staging for `nms` over host-resident frames of a fixed shape [*, N, 5+n_off]."""

    def __init__(self, N: int, n_off: int, chunk_frames: int = 2048, device="cuda:0"):
        self.N, self.P, self.chunk = int(N), 5 + int(n_off), int(chunk_frames)
        self.dev = torch.device(device)
        d = self.dev
        self.d_props = [torch.empty((self.chunk, self.N, self.P), dtype=torch.float32, device=d) for _ in range(2)]
        self.d_scores = [torch.empty((self.chunk, self.N), dtype=torch.float32, device=d) for _ in range(2)]
        self.d_keep = [torch.empty((self.chunk, self.N), dtype=torch.int64, device=d) for _ in range(2)]
        self.d_par = [torch.empty((self.chunk, self.N), dtype=torch.int64, device=d) for _ in range(2)]
        self.d_num = [torch.empty((self.chunk,), dtype=torch.int64, device=d) for _ in range(2)]
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(d) for _ in range(3))
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.launches = 0

    def alloc_outputs(self, F: int):
        """Pinned host tensors shaped like F reference calls: keep[F,N], num_to_keep[F], parent[F,N]."""
        return (torch.empty((F, self.N), dtype=torch.int64).pin_memory(),
                torch.empty((F,), dtype=torch.int64).pin_memory(),
                torch.empty((F, self.N), dtype=torch.int64).pin_memory())

    def __call__(self, props_h: torch.Tensor, scores_h: torch.Tensor, overlap, top_k, out=None, tuning=None, sync: bool = True):
        """Returns (keep, num_to_keep, parent) as pinned host tensors.  sync=True (default): the results are complete when the
        call returns.  sync=False: the device-to-host copies may still be in flight -- the current CUDA stream has been made to
        wait for them, so synchronise that stream (or record an event on it) before reading the tensors on the host."""
        if props_h.is_cuda or scores_h.is_cuda:
            raise RuntimeError("HostLaneNMS takes host tensors; call phnet_b200.ops.nms_batched for device tensors")
        if props_h.dtype != torch.float32 or props_h.dim() != 3 or tuple(props_h.shape[1:]) != (self.N, self.P):
            raise RuntimeError(f"props must be float32 [F, {self.N}, {self.P}]")
        F = props_h.shape[0]
        if out is None:
            out = self.alloc_outputs(F)
        keep_h, num_h, par_h = out
        cur = torch.cuda.current_stream(self.dev)
        for s in (self.s_in, self.s_run, self.s_out):
            s.wait_stream(cur)
        in_free = [None, None]     # event: compute on buffer b finished (its inputs may be overwritten)
        out_free = [None, None]    # event: D2H of buffer b finished (its outputs may be overwritten)
        for ci, f0 in enumerate(range(0, F, self.chunk)):
            f1 = min(F, f0 + self.chunk)
            n, b = f1 - f0, ci & 1
            with torch.cuda.stream(self.s_in):
                if in_free[b] is not None:
                    self.s_in.wait_event(in_free[b])
                self.d_props[b][:n].copy_(props_h[f0:f1], non_blocking=True)
                self.d_scores[b][:n].copy_(scores_h[f0:f1], non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(self.s_in)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(ev_in)
                if out_free[b] is not None:
                    self.s_run.wait_event(out_free[b])
                nms_batched(self.d_props[b][:n], self.d_scores[b][:n], overlap, top_k, tuning=tuning,
                            out=(self.d_keep[b][:n], self.d_num[b][:n], self.d_par[b][:n]))
                ev_run = torch.cuda.Event()
                ev_run.record(self.s_run)
                in_free[b] = ev_run
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_run)
                keep_h[f0:f1].copy_(self.d_keep[b][:n], non_blocking=True)
                num_h[f0:f1].copy_(self.d_num[b][:n], non_blocking=True)
                par_h[f0:f1].copy_(self.d_par[b][:n], non_blocking=True)
                ev_out = torch.cuda.Event()
                ev_out.record(self.s_out)
                out_free[b] = ev_out
            self.h2d_bytes += n * self.N * (self.P + 1) * 4
            self.d2h_bytes += n * (2 * self.N + 1) * 8
            self.launches += 1
        cur.wait_stream(self.s_out)
        cur.wait_stream(self.s_run)
        if sync:
            self.s_out.synchronize()
        return keep_h, num_h, par_h


def nms_host(props_h: torch.Tensor, scores_h: torch.Tensor, overlap, top_k, device="cuda:0", chunk_frames: int = 2048):
    """One-shot convenience wrapper around HostLaneNMS; synchronises before returning the host results."""
    if props_h.dim() == 2:
        k, n, p = nms_host(props_h[None], scores_h[None], overlap, top_k, device, 1)
        return [k[0], n[0], p[0]]
    pipe = HostLaneNMS(props_h.shape[1], props_h.shape[2] - 5, min(chunk_frames, max(1, props_h.shape[0])), device)
    return pipe(props_h, scores_h, overlap, top_k, sync=True)

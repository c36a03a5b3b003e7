"""Collection of the kept-lane results across GPUs through peer memory (NVLink / NVSwitch), without a collective.

Frames are sharded over ranks (one process per GPU, `phnet_b200.sharding`); the only exchange of the whole path is the
final collection of the kept lanes.  Instead of an all-gather that follows the kernel, every rank's NMS kernel stores the
compact record of each frame, int64[top_k + 1] = {kept indices, count}, straight into the result buffer of EVERY rank:
the buffers are dedicated allocations shared through CUDA IPC (C ABI `phnms_peer_*`, include/phnms.h), so the stores are
ordinary global stores that travel over NVLink while the kernel is still running.  A tiny flag kernel
(`phnms_peer_sync`) tells the peers that a step's records are complete.

The reference never gathers (it shards videos with a DistributedSampler and writes files per rank,
testOLV3.py:33-40,109-110); the collection exists for callers that want every rank to see all kept lanes.
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from . import _capi


class _RawDeviceMemory:
    """Lets torch view a raw device allocation (zero copy) through the CUDA array interface."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes // 8,), "typestr": "<i8", "data": (int(ptr), False), "version": 2}


class PeerCollector:
    """`nbuf` result buffers of [world * rows_per_rank, width] int64 on every rank, each mapped into every other rank.

    step protocol (all on the caller's stream):
        c = pc.collect_arg(i % nbuf)            # pass to nms_batched(..., collect=c): rank r writes rows r*rows .. of all ranks
        pc.signal(epoch)                         # after the kernel: my records of this step are complete everywhere
        pc.wait(epoch)                           # before reading pc.gathered(i % nbuf): everyone's records have arrived
    A buffer may be written again only when every rank has finished reading it.  With a consumer that runs `lag` steps
    behind (wait for epoch e - lag at step e, as bench.py does) a rank is at most lag + 1 steps ahead of the slowest one,
    so nbuf = 2 * lag + 3 rotating buffers are safe; with signal_and_wait every step (lock step), nbuf = 2 suffices.
    """

    def __init__(self, rows_per_rank: int, width: int, nbuf: int = 3, group=None, device=None):
        if not dist.is_initialized():
            raise RuntimeError("PeerCollector needs an initialised torch.distributed process group")
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > _capi.MAX_DST:
            raise RuntimeError(f"at most {_capi.MAX_DST} ranks")
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.rows, self.width, self.nbuf = int(rows_per_rank), int(width), int(nbuf)
        self.buf_bytes = (self.world * self.rows * self.width * 8 + 255) // 256 * 256
        self.flag_off = self.nbuf * self.buf_bytes
        self.nbytes = self.flag_off + 256                     # flags: one uint64 per rank, then an int status word
        L = _capi.lib()
        # Every rank walks through the same collectives whatever happens locally: a rank that cannot allocate or map
        # publishes the failure instead of leaving the others waiting in a collective, and then ALL ranks raise.
        self.local, self.base, err = 0, [], None
        with torch.cuda.device(self.dev):
            handle = ctypes.create_string_buffer(_capi.IPC_HANDLE_BYTES)
            try:
                ptr = ctypes.c_void_p()
                _capi.check(L.phnms_peer_alloc(self.nbytes, ctypes.byref(ptr), handle))
                self.local = int(ptr.value)
            except Exception as e:   # noqa: BLE001
                err = f"rank {self.rank}: peer_alloc: {e}"
            handles = [None] * self.world
            dist.all_gather_object(handles, None if err else bytes(handle.raw), group=group)
            if err is None and any(h is None for h in handles):
                err = "a peer could not allocate its buffer"
            if err is None:
                try:
                    for r in range(self.world):
                        if r == self.rank:
                            self.base.append(self.local)
                            continue
                        q = ctypes.c_void_p()
                        _capi.check(L.phnms_peer_open(handles[r], ctypes.byref(q)))
                        self.base.append(int(q.value))
                except Exception as e:   # noqa: BLE001
                    err = f"rank {self.rank}: peer_open: {e}"
            errs = [None] * self.world
            dist.all_gather_object(errs, err, group=group)
            if any(e is not None for e in errs):
                for r, b in enumerate(self.base):
                    if r != self.rank:
                        L.phnms_peer_close(b)
                if self.local:
                    L.phnms_peer_free(self.local)
                raise RuntimeError("peer memory unavailable: " + "; ".join(e for e in errs if e is not None))
        self._signal_dst = (ctypes.c_void_p * self.world)(*[b + self.flag_off + 8 * self.rank for b in self.base])
        self._mem = torch.as_tensor(_RawDeviceMemory(self.local, self.nbytes), device=self.dev)
        self.closed = False
        dist.barrier(group=group)     # every rank has mapped every buffer before anyone stores into them

    # -- views ---------------------------------------------------------------------------------------------------
    def gathered(self, buf: int) -> torch.Tensor:
        """[world * rows_per_rank, width] int64 view of result buffer `buf` on this rank (rows in rank order)."""
        n = self.world * self.rows * self.width
        o = buf * self.buf_bytes // 8
        return self._mem[o:o + n].view(self.world * self.rows, self.width)

    def status(self) -> int:
        """0, or 1 + the rank whose signal a wait gave up on (timeout)."""
        return int(self._mem[(self.flag_off + 8 * _capi.MAX_DST) // 8].item()) & 0xffffffff

    def collect_arg(self, buf: int, signal_epoch: int = 0, wait_epoch: int = 0, timeout_s: float = 10.0) -> _capi.Collect:
        """Descriptor for nms_batched(..., collect=...).  With signal_epoch / wait_epoch the record kernel itself completes
        the step across GPUs (its last block releases this rank's flag in every peer, then waits for wait_epoch of all
        ranks): one launch per step instead of records + phnms_peer_sync."""
        c = _capi.collect([b + buf * self.buf_bytes for b in self.base], rows=self.world * self.rows, width=self.width,
                          row0=self.rank * self.rows)
        if signal_epoch or wait_epoch:
            c.signal_epoch, c.wait_epoch, c.timeout_ns = int(signal_epoch), int(wait_epoch), int(timeout_s * 1e9)
            for r, b in enumerate(self.base):
                c.signal_dst[r] = b + self.flag_off + 8 * self.rank
            c.wait_src = self.local + self.flag_off
            c.status = self.local + self.flag_off + 8 * _capi.MAX_DST
            c.sync_counter = self.local + self.flag_off + 8 * _capi.MAX_DST + 64
        return c

    # -- completion flags ----------------------------------------------------------------------------------------
    def _sync(self, signal_epoch: int, wait_epoch: int, timeout_s: float):
        stream = torch.cuda.current_stream(self.dev).cuda_stream
        _capi.check(_capi.lib().phnms_peer_sync(self._signal_dst, self.local + self.flag_off, self.world, int(signal_epoch),
                                                int(wait_epoch), int(timeout_s * 1e9),
                                                self.local + self.flag_off + 8 * _capi.MAX_DST, stream))

    def signal(self, epoch: int):
        self._sync(epoch, 0, 0.0)

    def wait(self, epoch: int, timeout_s: float = 10.0):
        self._sync(0, epoch, timeout_s)

    def signal_and_wait(self, epoch: int, timeout_s: float = 10.0):
        self._sync(epoch, epoch, timeout_s)

    def close(self):
        if self.closed:
            return
        self.closed = True
        torch.cuda.synchronize(self.dev)
        dist.barrier(group=self.group)    # nobody is still storing into a buffer that is about to go away
        L = _capi.lib()
        self._mem = None
        with torch.cuda.device(self.dev):
            for r, b in enumerate(self.base):
                if r != self.rank:
                    L.phnms_peer_close(b)
            L.phnms_peer_free(self.local)


def local_collect(tensors, row0: int = 0) -> _capi.Collect:
    """Collection descriptor over plain tensors of this process (tests, single-GPU use): each [rows, top_k + 1] int64."""
    for t in tensors:
        if not (t.is_cuda and t.dtype == torch.int64 and t.is_contiguous() and t.dim() == 2):
            raise RuntimeError("collection buffers must be contiguous int64 CUDA tensors [rows, top_k + 1]")
    if len({tuple(t.shape) for t in tensors}) != 1:
        raise RuntimeError("collection buffers must all have the same shape")
    return _capi.collect([t.data_ptr() for t in tensors], rows=tensors[0].shape[0], width=tensors[0].shape[1], row0=row0)

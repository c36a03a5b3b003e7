#!/usr/bin/env python
"""bench.py -- frames/s of lane NMS (1000 proposals x 72 offsets, overlap 50, top_k 4) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W] # the reference semantics on host cores
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      # one rank per GPU, frames sharded (weak scaling)

One "step" = one pass of the hot path over one batch of synthetic frames (default 16384 frames per GPU = 5.05 GB of
proposals, far larger than the 126 MB L2, so every step streams from HBM).  Prints ONE JSON line (rank 0).

  value     frames/s, whole job, inputs resident in HBM, CUDA events around exactly K steps, max over ranks
  e2e       frames/s through the host-buffer API (phnet_b200.ops.HostLaneNMS): pinned host inputs, H2D, kernel, D2H of
            the reference-shaped results inside the timed region
  roofline  algorithmic bytes per launch / mean kernel duration (CUDA events around each launch) vs the measured HBM peak
  cpu_baseline  the CPU oracle (a port of the reference algorithm) on the host cores, bounded sample, rank 0, N=1 only
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "lane_nms_frames_per_sec"
UNIT = "frames/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=16384, help="frames per GPU per step")
    ap.add_argument("--e2e-frames", type=int, default=4096, help="frames per GPU per end-to-end step")
    ap.add_argument("--proposals", type=int, default=1000)
    ap.add_argument("--offsets", type=int, default=72)
    ap.add_argument("--top-k", type=int, default=4)
    ap.add_argument("--overlap", type=float, default=50.0)
    ap.add_argument("--cluster", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--max-clusters", type=int, default=0)
    ap.add_argument("--path", type=int, default=0)
    ap.add_argument("--collect", default="peer", choices=["peer", "nccl", "none"], help="N > 1: how the kept lanes are collected")
    ap.add_argument("--lag", type=int, default=2, help="peer collection: a step waits for the records of step e - lag of all ranks")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--groups", type=int, default=8, help="lane groups per synthetic frame (2-4 = road-like; PHNet max_lanes is 4)")
    ap.add_argument("--outlier-frac", type=float, default=0.1, help="fraction of proposals that belong to no lane group")
    ap.add_argument("--variant", type=int, default=0, help="0 auto, 2 register-resident cluster kernel, 3 streaming path, 4 one-launch small-frame kernel")
    ap.add_argument("--no-ref-cuda", action="store_true", help="skip timing the reference's own CUDA op (oracle/_ref)")
    return ap.parse_args()


def algorithmic_bytes_per_frame(N: int, n_off: int) -> int:
    """read proposals + scores once, write keep / parent / count once."""
    return N * (4 * n_off + 40) + 8


def workload_name(a) -> str:
    return (f"lane NMS, {a.proposals} proposals x {a.offsets} offsets fp32 per frame, overlap {a.overlap:g}, top_k {a.top_k} "
            f"(BASELINE configs[1] frame shape; batch scaled to {a.frames} frames/GPU/step so inputs exceed L2; "
            f"generator: {a.groups} lane groups per frame, {a.outlier_frac:g} outliers)")


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, torch_index: int):
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self._stop = threading.Event()
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(torch_index).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
            self.nv = pynvml
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:  # NVML missing: report that instead of inventing numbers
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for b, name in self.REASONS.items():
                    if bits & b and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.ok:
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self.ok:
            self.t.join()

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML sampling unavailable"}
        return {"sm_mhz": int(statistics.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(N, n_off, top_k, frames):
    """DRAM bytes per call from the committed `ncu --set full` captures of the same kernels, scaled per frame
    (profiles/roofline_traffic.json).  STATIC: a hardware counter cannot be read inside this run, so the figure is the
    capture's, not this run's -- the line says so in `traffic_source`."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            t = json.load(f)
        e = t[f"N{N}_No{n_off}_k{top_k}"]
        return float(e["dram_bytes_per_frame"]) * frames, "static: " + e["source"]
    except Exception:
        return None, "no committed capture for this shape"


def ref_cuda_op_rate(a, dev, frames=256):
    """frames/s of the reference's OWN CUDA op (libs/ops/csrc/nms_kernel.cu:147-192 via oracle/_ref), called per frame the
    way get_lanes calls it (libs/models/Router4OL.py:460-465), including the `keep[:num_to_keep]` host sync.  None when
    oracle/_ref was not built (it needs /root/reference at build time)."""
    try:
        import torch
        from oracle import ref_op
        from phnet_b200 import synth
        if ref_op.path(a.offsets) is None:
            return None
        props, scores = synth.make_frames(frames, a.proposals, a.offsets, seed=a.seed + 7, groups=a.groups, outlier_frac=a.outlier_frac)
        p, s = props.to(dev), scores.to(dev)
        for i in range(8):
            k, n, _ = ref_op.nms(p[i], s[i], a.overlap, a.top_k)
            _ = k[:n]
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for i in range(frames):
            k, n, _ = ref_op.nms(p[i], s[i], a.overlap, a.top_k)
            _ = k[:n]
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        return {"value": frames / dt, "unit": UNIT, "kind": "reference CUDA op (oracle/_ref), one call per frame + keep[:num] sync",
                "sample": f"{frames} frames in {dt:.2f} s"}
    except Exception as e:   # noqa: BLE001 -- a reported extra, never fatal
        return {"unavailable": f"{type(e).__name__}: {e}"}


# ---------------------------------------------------------------------------------------------------------------
def cpu_oracle_rate(a, frames, threads, repeat=1):
    """frames/s of the CPU oracle (literal reference algorithm: full 64x64-tile bitmask + serial collect)."""
    from oracle import oracle
    from phnet_b200 import synth
    props, scores = synth.make_frames(frames, a.proposals, a.offsets, seed=a.seed + 991, groups=a.groups, outlier_frac=a.outlier_frac)
    p, s = props.numpy(), scores.numpy()
    best = None
    for _ in range(repeat):
        t0 = time.perf_counter()
        oracle.nms_batched(p, s, None, a.overlap, a.top_k, lazy=False, threads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return frames / best, best


def run_reference(a):
    """`--impl reference`: the reference has no CPU implementation of this path (libs/ops is CUDA only, nms.cpp:40), so
    this arm times the CPU oracle -- the restatement of the reference algorithm pinned against the reference's own
    kernels -- on every host core.  Each step is a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    from phnet_b200 import synth
    cores = oracle.max_threads()
    probe_rate, _ = cpu_oracle_rate(a, max(cores, 8), cores)
    sample = max(cores, int(probe_rate * 1.0))                   # about one second of host work per step
    props, scores = synth.make_frames(sample, a.proposals, a.offsets, seed=a.seed, groups=a.groups, outlier_frac=a.outlier_frac)
    p, s = props.numpy(), scores.numpy()
    for _ in range(a.warmup):
        oracle.nms_batched(p, s, None, a.overlap, a.top_k, lazy=False, threads=cores)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        oracle.nms_batched(p, s, None, a.overlap, a.top_k, lazy=False, threads=cores)
    dt = time.perf_counter() - t0
    value = sample * a.steps / dt
    sample_desc = f"{sample} frames of the same workload per step (CPU-generated, seed {a.seed}), {cores} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "frames_per_step": sample, "proposals": a.proposals,
                   "offsets": a.offsets, "overlap": a.overlap, "top_k": a.top_k,
                   "note": "reference libs/ops has no CPU path; this is its algorithm restated in C (oracle/), all host cores"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample_desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist

    from phnet_b200 import _capi, sharding, synth
    from phnet_b200.ops import HostLaneNMS, nms_batched

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the lane-NMS op has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes its version banner to stdout when the first communicator comes up; rank 0 must print ONE JSON line,
        # so fd 1 points at stderr until the result is ready
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    _capi.lib()  # fail loudly before allocating anything if the native library is missing

    N, n_off, F = a.proposals, a.offsets, a.frames
    tune = _capi.tuning(path=a.path, cluster=a.cluster, threads=a.threads, max_clusters=a.max_clusters, variant=a.variant)
    plan = _capi.plan(F, N, n_off, tune, a.top_k)

    # synthetic frames, generated on the device rank by rank (weak scaling: every rank owns `F` frames)
    props, scores = synth.make_frames_chunked(F, N, n_off, seed=a.seed * 1000 + rank, device=dev, groups=a.groups,
                                              outlier_frac=a.outlier_frac)
    outs = [(torch.empty((F, N), dtype=torch.int64, device=dev), torch.empty((F,), dtype=torch.int64, device=dev),
             torch.empty((F, N), dtype=torch.int64, device=dev)) for _ in range(2)]
    gathered = [None]

    # Final collection of the kept lanes (the path's only exchange).  Preferred: the NMS kernel itself stores every frame's
    # compact record into the result buffer of EVERY rank through peer memory (NVLink / NVSwitch, phnet_b200/peer.py); a
    # one-warp flag kernel per step signals completion and waits for the previous step of all ranks.  If peer memory
    # cannot be set up on this box: one NCCL all-gather per step on the compute stream (a side stream would take an SM
    # away from the persistent NMS kernel).
    collector, collection = None, "none (single GPU)"
    lag = max(1, a.lag)
    nbuf = 2 * lag + 3
    mode = a.collect if world > 1 else "none"
    if world > 1 and mode == "none":
        collection = "none (diagnostic run: results stay on their rank)"
    elif world > 1 and mode != "nccl":
        try:
            from phnet_b200 import peer
            collector = peer.PeerCollector(F, a.top_k + 1, nbuf=nbuf)
            collection = (f"records stored into every rank's buffer over peer memory (CUDA IPC + NVLink) by the op's record kernel, whose "
                          f"last block also signals this step and waits for step e-{lag} of all ranks ({nbuf} rotating buffers): "
                          f"one extra launch per step")
        except Exception as e:   # noqa: BLE001 -- report why, fall back to the collective
            mode = "nccl"
            collection = f"NCCL all_gather_into_tensor per step (peer memory unavailable: {type(e).__name__}: {e})"
    elif world > 1:
        collection = "NCCL all_gather_into_tensor per step (requested)"
    step_no = [0]

    def step(i, ev_pair=None):
        b = i & 1
        cur = torch.cuda.current_stream(dev)
        step_no[0] += 1
        e = step_no[0]
        if ev_pair is not None:
            ev_pair[0].record(cur)
        if collector is not None:
            # Records of step e go to every rank; the same launch signals epoch e and waits until epoch e-lag of all ranks is
            # complete (a consumer of the gathered results runs `lag` steps behind the producer, which absorbs the
            # step-to-step jitter between GPUs).  A rank can run at most lag + 1 steps ahead of the slowest one and a
            # consumer reads records that are lag + 1 steps old, so 2 * lag + 3 rotating buffers guarantee that nobody
            # overwrites records a peer has yet to read.
            nms_batched(props, scores, a.overlap, a.top_k, tuning=tune, out=outs[b],
                        collect=collector.collect_arg(e % nbuf, signal_epoch=e, wait_epoch=max(e - lag, 0)))
        else:
            nms_batched(props, scores, a.overlap, a.top_k, tuning=tune, out=outs[b])
        if ev_pair is not None:
            ev_pair[1].record(cur)
        if collector is not None:
            if ev_pair is not None and len(ev_pair) > 2:
                ev_pair[2].record(cur)
            if e > lag:
                gathered[0] = collector.gathered((e - lag) % nbuf)
        elif mode == "nccl":
            packed = sharding.pack_kept(outs[b][0], outs[b][1], a.top_k)
            gathered[0] = sharding.gather_kept(packed, F * world)

    def fence():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for i in range(max(a.warmup, 3)):
        step(i)
    pairs = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in range(a.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # Everything slow and rank-dependent happens BEFORE the barrier that opens the timed region: NVML initialisation is
    # serialised across the processes of a box (measured: ranks entered the timed loop up to 12 ms apart when it sat
    # after the barrier, which a lock-step collection then pays as idle time in every rank's clock).
    clocks = ClockSampler(local)
    fence()
    with clocks:
        e0.record()
        for i in range(a.steps):
            step(i, pairs[i])
        e1.record()
        fence()
    ms_total = e0.elapsed_time(e1)
    kern_all = sorted(p[0].elapsed_time(p[1]) for p in pairs)
    kern_ms = statistics.mean(kern_all)
    kern_median, kern_min = statistics.median(kern_all), kern_all[0]
    sync_ms = statistics.mean(p[1].elapsed_time(p[2]) for p in pairs) if collector is not None else 0.0
    gap_ms = statistics.mean(pairs[i][1].elapsed_time(pairs[i + 1][0]) for i in range(a.steps - 1)) if a.steps > 1 else 0.0
    per_rank = None
    if world > 1:
        mine_t = torch.tensor([ms_total / a.steps, kern_ms, gap_ms, sync_ms], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine_t) for _ in range(world)]
        dist.all_gather(allr, mine_t)
        per_rank = {"ms_per_step": [round(float(t[0]), 4) for t in allr], "op_ms": [round(float(t[1]), 4) for t in allr],
                    "between_ops_ms": [round(float(t[2]), 4) for t in allr], "flag_kernel_ms": [round(float(t[3]), 4) for t in allr]}
    if collector is not None:
        # the last step's records: wait for them, then check the gathered buffer against this rank's own results
        collector.wait(step_no[0])
        torch.cuda.synchronize(dev)
        if collector.status() != 0:
            raise SystemExit(f"bench.py: peer collection timed out waiting for rank {collector.status() - 1}")
        last = collector.gathered(step_no[0] % nbuf)
        mine = sharding.pack_kept(outs[(a.steps - 1) & 1][0], outs[(a.steps - 1) & 1][1], a.top_k)
        assert torch.equal(last[rank * F:(rank + 1) * F], mine), "collected records differ from this rank's keep / num"
        assert bool((last[:, a.top_k] >= 1).all()), "a rank's records are missing from the gathered buffer"
        # every OTHER rank's records, too: one NCCL all-gather of the same packed results (outside the timed region) must
        # reproduce the peer-gathered buffer on every rank
        via_nccl = sharding.gather_kept(mine, F * world)
        assert torch.equal(via_nccl, last), f"rank {rank}: peer-memory collection differs from the NCCL all-gather of the same step"
        verified = "peer-gathered buffer == NCCL all-gather of every rank's records, on every rank"
    elif mode == "nccl":
        verified = "NCCL all-gather (the collection itself)"
    else:
        verified = None
    if rank == 0 and not a.no_cpu_baseline:
        # and the results themselves against the oracle, on a sample of this rank's frames
        from oracle import oracle as _oracle
        idx = torch.arange(0, F, max(1, F // 48))[:48]
        wk, wn, wp = _oracle.nms_batched(props[idx].cpu().numpy(), scores[idx].cpu().numpy(), None, a.overlap, a.top_k)
        o = outs[(a.steps - 1) & 1]
        assert (o[0][idx].cpu().numpy() == wk).all() and (o[1][idx].cpu().numpy() == wn).all() and (o[2][idx].cpu().numpy() == wp).all(), \
            "bench results differ from the CPU oracle"
    if world > 1:
        t = torch.tensor([ms_total, kern_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, kern_ms = float(t[0]), float(t[1])
    value = world * F * a.steps / (ms_total * 1e-3)

    # ---- end to end through the host-buffer API ------------------------------------------------------------
    e2e = None
    if not a.no_e2e:
        Fe = min(a.e2e_frames, F)
        pipe = HostLaneNMS(N, n_off, chunk_frames=min(1024, Fe), device=dev)
        props_h = torch.empty((Fe, N, 5 + n_off), dtype=torch.float32).pin_memory()
        scores_h = torch.empty((Fe, N), dtype=torch.float32).pin_memory()
        props_h.copy_(props[:Fe])
        scores_h.copy_(scores[:Fe])
        out_h = pipe.alloc_outputs(Fe)
        for _ in range(3):
            pipe(props_h, scores_h, a.overlap, a.top_k, out=out_h, tuning=tune)
        fence()
        pipe.h2d_bytes = pipe.d2h_bytes = pipe.launches = 0
        e2e_steps = max(3, min(a.steps, 20))
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(e2e_steps):     # (sync=False: steps overlap; the event below waits for the last device-to-host copy)
            pipe(props_h, scores_h, a.overlap, a.top_k, out=out_h, tuning=tune, sync=False)
        s1.record()
        fence()
        ms_e2e = s0.elapsed_time(s1)
        if world > 1:
            t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e2e = float(t[0])
        # the device path and the host path must agree on the same frames
        assert torch.equal(out_h[0], outs[(a.steps - 1) & 1][0][:Fe].cpu()), "e2e keep differs from device-resident run"
        # What the host side can deliver at most: the same bytes copied pinned host -> device with nothing else running, all
        # ranks at once (the ceiling for `e2e`: its steps move h2d_bytes_per_step over the same path).
        probe = torch.empty_like(props[:Fe])
        for _ in range(2):
            probe.copy_(props_h, non_blocking=True)
        fence()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        for _ in range(5):
            probe.copy_(props_h, non_blocking=True)
        q1.record()
        fence()
        ms_probe = q0.elapsed_time(q1) / 5
        if world > 1:
            t = torch.tensor([ms_probe], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_probe = float(t[0])
        h2d_gbs = props_h.numel() * 4 / (ms_probe * 1e-3) / 1e9
        del probe
        e2e = {"value": world * Fe * e2e_steps / (ms_e2e * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": pipe.h2d_bytes // e2e_steps, "d2h_bytes_per_step": pipe.d2h_bytes // e2e_steps,
               "frames_per_step_per_gpu": Fe, "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps,
               "h2d_gbs_per_gpu": (pipe.h2d_bytes // e2e_steps) / (ms_e2e / e2e_steps * 1e-3) / 1e9,
               "h2d_ceiling_gbs_per_gpu": h2d_gbs,
               "h2d_ceiling_note": "pinned host -> device copy of the same frames alone, all ranks at once, slowest rank: what the host / PCIe side delivers at most",
               "api": "phnet_b200.ops.HostLaneNMS (pinned host tensors in, reference-shaped keep/num/parent out)"}

    # ---- CPU baseline: the oracle on the host cores, bounded sample ------------------------------------------
    cpu = None
    ref_cuda = None
    if rank == 0 and world == 1 and not a.no_ref_cuda:
        ref_cuda = ref_cuda_op_rate(a, dev)
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        from oracle import oracle
        cores = oracle.max_threads()
        probe, _ = cpu_oracle_rate(a, max(cores, 8), cores)
        sample = max(cores, int(probe * 12.0))     # about 12 s of host work
        rate, dt = cpu_oracle_rate(a, sample, cores)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{sample} frames of the same workload in {dt:.1f} s on {cores} threads (oracle/lane_nms_oracle.c, literal N^2 form)"}

    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        bpf = algorithmic_bytes_per_frame(N, n_off)
        achieved = F * bpf / (kern_ms * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic(N, n_off, a.top_k, F)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "frames_per_gpu_per_step": F, "global_frames_per_step": F * world,
                       "proposals": N, "offsets": n_off, "overlap": a.overlap, "top_k": a.top_k,
                       "groups": a.groups, "outlier_frac": a.outlier_frac,
                       "l2": f"inputs are {F * N * (6 + n_off) * 4 / 1e9:.2f} GB per GPU per step, larger than the 126 MB L2 (no flush needed)",
                       "parallelism": f"frames sharded x{world}; no data-path collective; kept lanes collected on every rank once per step" if world > 1 else "single GPU",
                       "collection": collection, "collection_verified": verified, "per_rank": per_rank,
                       "plan": plan},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_frame": bpf,
                         "kernel": ("phnms_select_kernel + phnms_stream_kernel + resume pass (one C-ABI call; the stream kernel is ~90 % of it, profiles/)" if plan.get("variant") == 3
                                    else "phnms_small_kernel (one launch: persistent CTAs, one frame at a time, rows in registers)" if plan.get("variant") == 4
                                    else "phnms_topm_kernel + phnms_freg_kernel (one C-ABI call)" if plan.get("variant") == 2
                                    else "phnms_fused_kernel" if plan["path"] == 1 else "phnms_order/mask/scan kernels"),
                         "kernel_ms_per_launch": kern_ms, "kernel_ms_median": kern_median, "kernel_ms_min": kern_min,
                         "frac_median": F * bpf / (kern_median * 1e-3) / 1e9 / peak, "frac_best": F * bpf / (kern_min * 1e-3) / 1e9 / peak},
            "clocks": clocks.summary(),
            "gpu_launches": a.steps * (plan["launches"] + (1 if collector is not None else 0)),
        }
        if e2e is not None:
            line["e2e"] = e2e
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if ref_cuda is not None:
            line["ref_cuda_op"] = ref_cuda
        if world > 1:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if collector is not None:
        collector.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)

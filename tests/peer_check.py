"""Run under torchrun with N >= 2 ranks: every rank runs lane NMS on its shard and stores the compact kept-lane records
into EVERY rank's buffer through peer memory (phnet_b200.peer); the gathered result must equal the single-process answer
of the CPU oracle, and agree with the NCCL all-gather path.  Prints timing of both collection methods."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle  # noqa: E402
from phnet_b200 import peer, sharding, synth  # noqa: E402
from phnet_b200.ops import nms_batched  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    N, n_off, top_k, Fr = 1000, 72, 4, 64
    F = Fr * world
    props, scores = synth.make_frames(F, N, n_off, seed=5)            # same seed everywhere; each rank takes its block
    f0, f1 = sharding.shard_range(F, rank, world)
    p, s = props[f0:f1].to(dev), scores[f0:f1].to(dev)
    pc = peer.PeerCollector(Fr, top_k + 1, nbuf=3)
    wk, wn, _ = oracle.nms_batched(props.numpy(), scores.numpy(), None, 50.0, top_k)
    want = sharding.pack_kept(torch.from_numpy(wk), torch.from_numpy(wn), top_k)
    for step in range(1, 9):                                           # several epochs over the three buffers
        b = step % 3
        pc.gathered(b).fill_(-1)
        torch.cuda.synchronize()
        dist.barrier()
        if step % 2:      # records, then the flag kernel
            keep, num, _ = nms_batched(p, s, 50.0, top_k, collect=pc.collect_arg(b))
            pc.signal_and_wait(step)
        else:             # records + completion in one launch
            keep, num, _ = nms_batched(p, s, 50.0, top_k, collect=pc.collect_arg(b, signal_epoch=step, wait_epoch=step))
        got = pc.gathered(b).cpu()
        assert pc.status() == 0, f"rank {rank}: wait timed out on rank {pc.status() - 1}"
        assert torch.equal(got, want), f"rank {rank} step {step}: gathered records differ from the oracle"
        ref = sharding.gather_kept(sharding.pack_kept(keep, num, top_k), F).cpu()
        assert torch.equal(ref, want), "NCCL all-gather path differs"
    # timing of the two collection methods on the bench shape (kernel + collection per step, CUDA events)
    Fb = 4096
    pb, sb = synth.make_frames_chunked(Fb, N, n_off, seed=rank, device=dev)
    pcb = peer.PeerCollector(Fb, top_k + 1, nbuf=3)
    out = (torch.empty((Fb, N), dtype=torch.int64, device=dev), torch.empty((Fb,), dtype=torch.int64, device=dev),
           torch.empty((Fb, N), dtype=torch.int64, device=dev))

    def timed(fn, reps=20):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(3, 3 + reps):
            fn(i)
        e1.record()
        torch.cuda.synchronize(); dist.barrier()
        return e0.elapsed_time(e1) / reps

    epoch = [100]

    def step_peer(i):
        epoch[0] += 1
        nms_batched(pb, sb, 50.0, top_k, out=out, collect=pcb.collect_arg(i % 3, signal_epoch=epoch[0], wait_epoch=epoch[0]))

    def step_nccl(i):
        nms_batched(pb, sb, 50.0, top_k, out=out)
        sharding.gather_kept(sharding.pack_kept(out[0], out[1], top_k), Fb * world)

    def step_none(i):
        nms_batched(pb, sb, 50.0, top_k, out=out)

    t_none, t_peer, t_nccl = timed(step_none), timed(step_peer), timed(step_nccl)
    assert pcb.status() == 0
    pcb.close()
    pc.close()
    if rank == 0:
        print(f"peer collection ok: world {world}; ms/step kernel only {t_none:.4f}, + peer-memory collection {t_peer:.4f}, "
              f"+ NCCL all-gather {t_nccl:.4f}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

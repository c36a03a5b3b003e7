"""Per-phase cycle breakdown of the register-resident fused kernel (CTA 0, thread 0), via phnms_forward_f32_trace."""
import ctypes, os, sys, collections
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phnet_b200 import _capi, synth

N = int(os.environ.get("N", 1000)); n_off = int(os.environ.get("NOFF", 72)); F = int(os.environ.get("F", 4736))
tune = _capi.tuning(path=1, cluster=int(os.environ.get("CLUSTER", 2)), threads=int(os.environ.get("THREADS", 512)), variant=2, max_clusters=int(os.environ.get("MAXC", 0)))
dev = torch.device("cuda:0")
props, scores = synth.make_frames_chunked(F, N, n_off, seed=0, device=dev)
keep = torch.empty((F, N), dtype=torch.int64, device=dev); num = torch.empty((F,), dtype=torch.int64, device=dev)
par = torch.empty((F, N), dtype=torch.int64, device=dev)
TL = 1 << 16
trace = torch.zeros(TL, dtype=torch.int64, device=dev)
L = _capi.lib()
wsb = L.phnms_workspace_bytes(F, N, n_off, ctypes.byref(tune))
ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=dev)
for it in range(3):
    trace.zero_()
    rc = L.phnms_forward_f32_trace(props.data_ptr(), scores.data_ptr(), None, F, N, n_off, 50.0, 4, 0, keep.data_ptr(),
                                   num.data_ptr(), par.data_ptr(), ws.data_ptr(), wsb, ctypes.byref(tune), torch.cuda.current_stream().cuda_stream,
                                   trace.data_ptr(), TL)
    assert rc == 0, rc
torch.cuda.synchronize()
t = trace.cpu().tolist()
ev = [(t[i], t[i + 1]) for i in range(0, TL - 1, 2) if t[i] != 0]
names = {1: "frame start", 2: "wait slab (+sync)", 3: "fill regs (+bitonic, sync)", 4: "prefetch next scores", 16: "request next slab (TMA issue)", 5: "warp top-M (+sync)",
         6: "CTA top-M", 7: "publish", 8: "exchange wait", 9: "merge (+sync)", 10: "spare-lane load", 11: "round tail", 14: "round set-up", 15: "round offset loop",
         12: "round barrier", 13: "outputs"}
tot = collections.defaultdict(int); cnt = collections.defaultdict(int)
frames = sum(1 for a, _ in ev if a == 1)
for (a0, c0), (a1, c1) in zip(ev[:-1], ev[1:]):
    if a1 == 1:
        tot["(between frames)"] += c1 - c0; cnt["(between frames)"] += 1
        continue
    tot[names[a1]] += c1 - c0; cnt[names[a1]] += 1
total = sum(tot.values())
print(f"plan {_capi.plan(F, N, n_off, tune)}")
print(f"frames traced {frames}; cycles per frame {total / max(frames,1):.0f}")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{k:32s} {v / frames:9.0f} cyc/frame  {100 * v / total:5.1f}%   x{cnt[k] / frames:.2f}/frame  avg {v / cnt[k]:.0f}")

"""Where the host time of one drop-in call goes (N = 240, one frame per call, async): indexing the clip tensor, the Python
wrapper, the pybind shim, and -- for scale -- the smallest torch op.  Prints microseconds per call."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phnet_b200 import _capi, synth
from phnet_b200.ops import nms

dev = torch.device("cuda:0")
F = 2000
out = {}
for n_off, N, top_k in ((72, 240, 4), (36, 240, 8)):
    props, scores = synth.make_frames(512, N, n_off, seed=3, groups=4)
    props, scores = props.to(dev), scores.to(dev)
    ps = [props[f % 512] for f in range(F)]
    ss = [scores[f % 512] for f in range(F)]
    sh = _capi.shim()
    one = torch.zeros(1, device=dev)

    def timed(fn):
        for f in range(50):
            fn(f)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for f in range(F):
            fn(f)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        return round((t1 - t0) / F * 1e6, 2), round((t2 - t0) / F * 1e6, 2)

    tag = f"N{N}_No{n_off}_k{top_k}"
    out[tag] = {
        "index_only": timed(lambda f: (props[f % 512], scores[f % 512])),
        "nms_indexing_inside": timed(lambda f: nms(props[f % 512], scores[f % 512], 50.0, top_k)),
        "nms_presliced": timed(lambda f: nms(ps[f], ss[f], 50.0, top_k)),
        "shim_direct": timed(lambda f: sh.nms_forward(ps[f], ss[f], 50.0, top_k)),
        "torch_add_": timed(lambda f: one.add_(1.0)),
        "torch_empty": timed(lambda f: torch.empty(481, dtype=torch.int64, device=dev)),
    }
print(json.dumps(out))

"""Loader for the LIVE reference op built by oracle/build_ref.py (test infrastructure, not product).

`load(n_off)` returns the pybind module compiled from the reference's own libs/ops/csrc sources
(`nms_forward(boxes, scores, thresh, top_k)`, libs/ops/csrc/nms.cpp:44-61) or None when it has not been built.
Needs a CUDA device to run (the reference op has no CPU path, nms.cpp:40).
"""
import importlib.util
import os

HERE = os.path.dirname(os.path.abspath(__file__))
_cache = {}


def path(n_off: int):
    p = os.path.join(HERE, "_ref", f"phnet_ref_nms_{n_off}.so")
    return p if os.path.exists(p) else None


def load(n_off: int):
    if n_off in _cache:
        return _cache[n_off]
    p = path(n_off)
    mod = None
    if p is not None:
        import torch  # noqa: F401  (libtorch symbols must be loaded first)
        name = f"phnet_ref_nms_{n_off}"
        spec = importlib.util.spec_from_file_location(name, p)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    _cache[n_off] = mod
    return mod


def nms(boxes, scores, overlap, top_k):
    """The reference's Python surface (libs/ops/nms.py:32-33) on top of the live build."""
    mod = load(boxes.shape[1] - 5)
    if mod is None:
        raise RuntimeError("reference op not built for this offset count (oracle/build_ref.py)")
    return mod.nms_forward(boxes, scores, overlap, top_k)

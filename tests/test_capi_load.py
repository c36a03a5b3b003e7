"""CPU checks of the drop-in boundary: the C-ABI library builds, loads, exports every symbol include/phnms.h declares,
plans launches sensibly and rejects bad arguments before touching a GPU (no compute calls here)."""
import ctypes
import os
import re

import pytest

from phnet_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "phnms.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(phnms_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = _capi.lib()
    syms = declared_symbols()
    assert set(syms) == set(_capi.EXPORTS), "include/phnms.h and the ctypes binding disagree"
    for s in syms:
        assert hasattr(L, s), f"libphnms.so does not export {s}"
    assert L.phnms_abi_version() == _capi.ABI_VERSION == 5


def test_library_is_sm100a_only_and_has_tma_and_cluster_code():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _capi.SO_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out
    sass = subprocess.run(["cuobjdump", "-sass", _capi.SO_PATH], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass          # cp.async.bulk (TMA 1-D bulk copy)
    assert "UCGABAR" in sass         # barrier.cluster
    # the fp32 predicate must never be contracted: no FFMA in any kernel that evaluates it
    # (expf in the get_lanes front-end legitimately uses FMAs)
    import re
    funcs = re.split(r"Function : ", sass)[1:]
    checked = 0
    for fn in funcs:
        name = fn.split("\n", 1)[0]
        if "get_lanes" in name:
            continue    # (its softmax is expf, FMA-based like ATen's; the predicate code in it is the select kernel's, checked below)
        if any(k in name for k in ("freg_kernel", "fused_kernel", "mask_kernel", "stream_kernel", "select_kernel")):
            assert "FFMA" not in fn, name
            checked += 1
            if "stream_kernel" in name:     # the streaming kernel: TMA bulk copies, cp.async on an mbarrier, packed fp32 subtraction
                assert "UBLKCP" in fn and "LDGSTS" in fn and "SYNCS" in fn, name
                if "ELi1EEE" not in name:
                    assert "FADD2" in fn, name
    assert checked >= 12


def test_error_strings_and_argument_checks_without_a_gpu():
    L = _capi.lib()
    assert b"offsets" in L.phnms_error_string(-2)
    assert b"MAX_COL_BLOCKS" in L.phnms_error_string(-3)
    nul = None
    # shape errors are detected before any CUDA call
    assert L.phnms_forward_f32(nul, nul, nul, 1, 10, 0, 50.0, 4, 0, nul, nul, nul, nul, 0, None, nul) == -2
    assert L.phnms_forward_f32(nul, nul, nul, 1, 10, 251, 50.0, 4, 0, nul, nul, nul, nul, 0, None, nul) == -2
    assert L.phnms_forward_f32(nul, nul, nul, 1, 64000, 72, 50.0, 4, 0, nul, nul, nul, nul, 0, None, nul) == -3
    assert L.phnms_forward_f32(nul, nul, nul, -1, 10, 72, 50.0, 4, 0, nul, nul, nul, nul, 0, None, nul) == -1
    assert L.phnms_forward_f32(nul, nul, nul, 1, 10, 72, 50.0, -4, 0, nul, nul, nul, nul, 0, None, nul) == -1
    assert L.phnms_forward_f32(nul, nul, nul, 1, 10, 72, 50.0, 4, 7, nul, nul, nul, nul, 0, None, nul) == -1
    assert L.phnms_forward_f32(nul, nul, nul, 0, 10, 72, 50.0, 4, 0, nul, nul, nul, nul, 0, None, nul) == 0   # F == 0
    assert L.phnms_forward_f32(nul, nul, nul, 1, 10, 72, 50.0, 4, 0, nul, nul, nul, nul, 0, None, nul) == -1  # null outputs
    assert L.phnms_order_f32(nul, nul, 1, 10, 0, nul, nul, 0, nul) == -1
    # collection: descriptor checks (no destination, too many, top_k == 0, misaligned buffer) come before any CUDA call
    args = (nul, nul, nul, 1, 10, 72, 50.0, 4, 0, nul, nul, nul, nul, 0, None, nul)
    assert L.phnms_forward_collect_f32(*args, None) == -1
    c = _capi.Collect()
    c.n_dst = 0
    assert L.phnms_forward_collect_f32(*args, ctypes.byref(c)) == -1
    c.n_dst = _capi.MAX_DST + 1
    assert L.phnms_forward_collect_f32(*args, ctypes.byref(c)) == -1
    c = _capi.collect([4096], rows=1, width=5)
    assert L.phnms_forward_collect_f32(*(args[:7] + (0,) + args[8:]), ctypes.byref(c)) == -1       # top_k == 0
    assert L.phnms_forward_collect_f32(*args, ctypes.byref(_capi.collect([4100], rows=1, width=5))) == -1   # not 8-byte aligned
    # a record is never stored outside a destination: width must be top_k + 1 and row0 + F must fit (checked before any launch;
    # with valid descriptors the same null-pointer call gets as far as the shape / pointer checks: -1 either way, so compare
    # against a descriptor that is fine and a shape that is not)
    bad_shape = args[:5] + (0,) + args[6:]       # n_off == 0 -> -2, but only if the descriptor passed
    assert L.phnms_forward_collect_f32(*bad_shape, ctypes.byref(_capi.collect([4096], rows=1, width=5))) == -2
    assert L.phnms_forward_collect_f32(*bad_shape, ctypes.byref(_capi.collect([4096], rows=1, width=4))) == -1   # width != top_k + 1
    assert L.phnms_forward_collect_f32(*bad_shape, ctypes.byref(_capi.collect([4096], rows=1, width=5, row0=1))) == -1   # row0 + F > rows
    assert L.phnms_forward_collect_f32(*bad_shape, ctypes.byref(_capi.collect([4096], rows=0, width=5))) == -1
    assert L.phnms_peer_sync(None, nul, 0, 1, 1, 0, nul, nul) == -1
    assert L.phnms_peer_sync(None, nul, 2, 0, 0, 0, nul, nul) == 0                                  # nothing to do
    assert L.phnms_peer_open(None, None) == -1 and L.phnms_peer_close(None) == -1 and L.phnms_peer_free(None) == -1
    with pytest.raises(_capi.PhnmsError):
        _capi.check(-6)


def test_plans():
    p = _capi.plan(16384, 1000, 72)             # the headline shape: select + stream (+ resume) kernels, one CTA per SM
    assert p["path"] == _capi.PATH_FUSED and p["variant"] == _capi.FUSED_STREAM and p["launches"] == 3
    assert p["cluster"] == 1 and p["threads"] == 512 and p["grid"] == 148 and p["smem_bytes"] <= 232448
    ws = 16384 * 16 * (32 + 4 * 80)             # per-frame block (capacity of the cluster kernel's 16 candidate slots)
    assert ws < p["workspace_bytes"] <= ws + 2 * 16384 * 4 + 2048
    for top_k in (0, 9, 1000):                  # outside [1, 8]: the register-resident cluster kernel
        q = _capi.plan(16384, 1000, 72, None, top_k)
        assert q["variant"] == _capi.FUSED_REG and q["cluster"] * q["rows_per_cta"] >= 1000 and q["launches"] == 2
        assert q["grid"] % q["cluster"] == 0 and q["workspace_bytes"] == p["workspace_bytes"]
    assert _capi.plan(16384, 1000, 72, None, 8)["variant"] == _capi.FUSED_STREAM
    assert _capi.plan(1, 240, 72, _capi.tuning(variant=_capi.FUSED_REG))["cluster"] == 1   # the real OpenLane-V shape fits one CTA
    assert _capi.plan(4, 100, 50)["variant"] == _capi.FUSED_SMEM        # other offset counts: shared-memory cluster kernel
    with pytest.raises(_capi.PhnmsError):
        _capi.plan(4, 100, 50, _capi.tuning(variant=_capi.FUSED_STREAM))
    with pytest.raises(_capi.PhnmsError):
        _capi.plan(4, 1000, 72, _capi.tuning(variant=_capi.FUSED_STREAM), 0)
    small = _capi.plan(8, 1000, 72)             # fewer frames than SMs: frames are cut into units so that every SM has work
    assert small["variant"] == _capi.FUSED_STREAM and small["grid"] > 8
    one = _capi.plan(1, 240, 72)                # PHNet's own call, one frame: launch-latency bound -> ONE launch, no workspace
    assert one["variant"] == _capi.FUSED_SMALL and one["launches"] == 1 and one["grid"] == 1 and one["workspace_bytes"] == 0
    mid = _capi.plan(1, 1000, 72)               # one frame of more than 512 proposals: the cluster kernel, still one launch
    assert mid["variant"] == _capi.FUSED_REG and mid["launches"] == 1
    assert _capi.plan(1, 8192, 72)["path"] == _capi.PATH_FUSED          # the whole stress sweep stays on the fused path
    big = _capi.plan(1, 40000, 72)
    assert big["path"] == _capi.PATH_TILED and big["launches"] == 3 and big["workspace_bytes"] > 0
    assert _capi.lib().phnms_workspace_bytes(1, 40000, 72, None) == big["workspace_bytes"]
    t = _capi.tuning(path=_capi.PATH_FUSED, cluster=2, threads=512)
    assert _capi.plan(8, 1000, 72, t)["cluster"] == 2
    with pytest.raises(_capi.PhnmsError):
        _capi.plan(8, 1000, 72, _capi.tuning(cluster=3))
    with pytest.raises(_capi.PhnmsError):
        _capi.plan(8, 8192, 72, _capi.tuning(path=_capi.PATH_FUSED, cluster=1))   # does not fit one CTA


def test_python_mirror_refuses_cpu_tensors():
    import torch
    from phnet_b200.ops import nms
    with pytest.raises(RuntimeError, match="CUDA"):
        nms(torch.zeros(8, 77), torch.zeros(8), overlap=50, top_k=4)


def test_install_as_libs_ops():
    import sys
    import phnet_b200
    phnet_b200.install_as_libs_ops()
    from libs.ops import nms as shim  # noqa: E402
    from phnet_b200.ops import nms
    assert shim is nms
    for k in ("libs", "libs.ops", "libs.ops.nms"):
        sys.modules.pop(k, None)


def test_header_is_plain_c(tmp_path):
    """include/phnms.h is the drop-in boundary for any host language: it must compile as C (and as C++) on its own, and a C
    program linked against libphnms.so must resolve the entry points."""
    import subprocess
    src = tmp_path / "use_phnms.c"
    src.write_text('#include "phnms.h"\n'
                   '#include <stdio.h>\n'
                   'int main(void) {\n'
                   '    phnms_plan pl; phnms_collect c; c.n_dst = 0; (void)c;\n'
                   '    int rc = phnms_plan_query(16, 1000, 72, 0, &pl);\n'
                   '    printf("%d %d %d %d %zu\\n", phnms_abi_version(), rc, pl.cluster, pl.threads, phnms_workspace_bytes(1, 40000, 72, 0));\n'
                   '    return phnms_forward_f32(0, 0, 0, 1, 10, 0, 50.f, 4, 0, 0, 0, 0, 0, 0, 0, 0) == PHNMS_ERR_N_OFFSETS ? 0 : 1;\n'
                   '}\n')
    inc = os.path.join(ROOT, "include")
    so_dir = os.path.dirname(_capi.SO_PATH)
    exe = tmp_path / "use_phnms"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", inc, str(src), "-o", str(exe), "-L", so_dir,
                    "-l:libphnms.so", f"-Wl,-rpath,{so_dir}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    ver, rc, cluster, threads, ws = out.stdout.split()
    assert int(ver) == _capi.ABI_VERSION and int(rc) == 0 and int(cluster) == 1 and int(threads) == 512 and int(ws) > 0
    subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I", inc, "-x", "c++", os.path.join(inc, "phnms.h")], check=True)


def test_pybind_shim_has_the_reference_native_surface():
    """phnet_b200/csrc/nms_impl.so is a pybind11 module named like the reference's (`nms_impl`) exporting `nms_forward(boxes,
    scores, thresh, top_k)` (libs/ops/csrc/nms.cpp:44-61); argument checks raise RuntimeError like CHECK_CUDA / CHECK_CONTIGUOUS."""
    import torch
    from phnet_b200 import build
    build.build_shim()
    sh = _capi.shim()
    assert sh is not None and sh.__name__ == "nms_impl" and sh.abi_version() == _capi.ABI_VERSION
    assert "nms_forward" in sh.nms_forward.__doc__ and "Tensor" in sh.nms_forward.__doc__
    with pytest.raises(RuntimeError, match="CUDA"):
        sh.nms_forward(torch.zeros(8, 77), torch.zeros(8), 50.0, 4)
    with pytest.raises(TypeError):
        sh.nms_forward(torch.zeros(8, 77), torch.zeros(8), 50.0, -1)      # unsigned long top_k (nms.cpp:48)


def _reference_style_package(tmp_path):
    """A package shaped like the reference's `libs/ops` -- `nms.py` does `from . import nms_impl` and forwards to
    `nms_impl.nms_forward(boxes, scores, overlap, top_k)` (libs/ops/nms.py:29-33) -- with THIS repo's nms_impl.so (and the
    libphnms.so it links against) dropped in where the reference's compiled module would be (option 0)."""
    from phnet_b200 import build
    build.build_shim()
    pkg = tmp_path / "refstyle_libs" / "ops"
    pkg.mkdir(parents=True)
    (tmp_path / "refstyle_libs" / "__init__.py").write_text("")
    (pkg / "__init__.py").write_text("from .nms import nms\n__all__ = ['nms']\n")
    (pkg / "nms.py").write_text("from . import nms_impl\n\n\ndef nms(boxes, scores, overlap, top_k):\n"
                                "    return nms_impl.nms_forward(boxes, scores, overlap, top_k)\n")
    # a COPY, as a maintainer would install it -- and a separate file is a separate dlopen instance: initialising one pybind11
    # module object twice in a process (a symlink to the file phnet_b200 itself loads) re-runs PyInit on the same static
    # PyModuleDef, which crashed the interpreter depending on test order
    import shutil
    shutil.copy(build.SHIM_SO, pkg / "nms_impl.so")
    os.symlink(build.SO, pkg / "libphnms.so")
    return str(tmp_path)


def test_shim_drops_into_a_reference_style_package(tmp_path):
    import importlib
    import sys
    import torch
    root = _reference_style_package(tmp_path)
    sys.path.insert(0, root)
    try:
        ops = importlib.import_module("refstyle_libs.ops")
        with pytest.raises(RuntimeError, match="CUDA"):
            ops.nms(torch.zeros(8, 77), torch.zeros(8), overlap=50, top_k=4)     # bound and checked like the reference's op
    finally:
        sys.path.remove(root)
        for k in [k for k in sys.modules if k.startswith("refstyle_libs")]:
            del sys.modules[k]

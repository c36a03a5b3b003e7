"""Build recipe for the native library: phnet_b200/csrc/libphnms.so (sm_100a only, in-tree).

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(CSRC, "libphnms.so")
SOURCES = ["phnms.cu"]
HEADERS = ["common.cuh", "fused_nms.cuh", "fused_reg.cuh", "topm.cuh", "select.cuh", "stream.cuh", "small.cuh", "assign.cuh", "tiled_nms.cuh", "frontend.cuh",
           os.path.join("..", "..", "include", "phnms.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",            # the reference's fp32 sequence has no FMA; never contract
    "-shared", "-Xcompiler", "-fPIC",
    "-cudart", "static",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libphnms.so")
    return exe


def stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, extra: list[str] | None = None) -> str:
    if not force and not stale():
        return SO
    # (written next to the target and renamed over it: a process that has the old library mapped -- pytest after a header edit --
    # keeps its old inode instead of seeing the file change under its feet)
    tmp_so = SO + f".tmp{os.getpid()}"
    cmd = [nvcc()] + NVCC_FLAGS + (extra or []) + ["-o", tmp_so] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        if os.path.exists(tmp_so):
            os.remove(tmp_so)
        raise RuntimeError("nvcc failed building libphnms.so")
    os.replace(tmp_so, SO)
    return SO


SHIM_SRC = os.path.join(CSRC, "nms_impl.cpp")
SHIM_SO = os.path.join(CSRC, "nms_impl.so")


def shim_stale() -> bool:
    if not os.path.exists(SHIM_SO):
        return True
    t = os.path.getmtime(SHIM_SO)
    return any(os.path.getmtime(d) > t for d in (SHIM_SRC, os.path.join(CSRC, "..", "..", "include", "phnms.h")))


def build_shim(force: bool = False, verbose: bool = False) -> str:
    """The pybind11 module `nms_impl` (phnet_b200/csrc/nms_impl.cpp): the reference's native surface `nms_forward(boxes, scores,
    thresh, top_k)` (libs/ops/csrc/nms.cpp:44-61) as a thin ATen adapter over libphnms.so.  Built in-tree with torch's
    cpp_extension (host compiler only, ~1 min); needs libphnms.so next to it at run time ($ORIGIN rpath)."""
    if not force and not shim_stale():
        return SHIM_SO
    import sys
    import tempfile
    if not os.path.exists(SO):
        build()
    tmp = tempfile.mkdtemp(prefix="phnms_shim_")
    # Built in a CHILD process: cpp_extension.load() also dlopens what it built, and a second copy of the module loaded into a
    # process that already holds one (pytest after phnet_b200 imported its shim) crashed the interpreter.
    code = (
        "import os, sys\n"
        "from torch.utils.cpp_extension import load\n"
        "os.environ.setdefault('MAX_JOBS', '4')\n"
        "try:\n"
        "    load(name='nms_impl', sources=[sys.argv[1]], extra_cflags=['-O2'], with_cuda=True,\n"
        "         extra_ldflags=['-L' + sys.argv[2], '-l:libphnms.so', '-Wl,-rpath,\\\\$$ORIGIN'],\n"
        "         build_directory=sys.argv[3], is_python_module=False, verbose=bool(int(sys.argv[4])))\n"
        "except OSError:\n"
        "    pass\n"   # built, but not loadable from the temporary directory: libphnms.so is found through $ORIGIN, i.e. next to it
    )
    try:
        res = subprocess.run([sys.executable, "-c", code, SHIM_SRC, CSRC, tmp, "1" if verbose else "0"], capture_output=not verbose, text=True)
        if not os.path.exists(os.path.join(tmp, "nms_impl.so")):
            if res.stdout or res.stderr:
                print(res.stdout, res.stderr)
            raise RuntimeError("building the nms_impl shim failed")
        shutil.copy(os.path.join(tmp, "nms_impl.so"), SHIM_SO + f".tmp{os.getpid()}")
        os.replace(SHIM_SO + f".tmp{os.getpid()}", SHIM_SO)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return SHIM_SO


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose="-v" in sys.argv))
    if "--shim" in sys.argv:
        print(build_shim(force=True, verbose="-v" in sys.argv))

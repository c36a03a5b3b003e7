"""Build recipe for the LIVE reference op (test infrastructure, not product).

Compiles the reference's own lane-NMS sources *where they lie* under
/root/reference/libs/ops/csrc (nms.cpp, nms_kernel.cu) for sm_100a into
oracle/_ref/ (git-ignored, NOT gpurun-ignored: the .so travels to the GPU box).

No reference source is copied into the repo.  Two edits are needed to build it
against torch 2.11 / for 72 offsets; they are applied by streaming the file
through a text substitution into a temp dir OUTSIDE the repo:

  * nms_kernel.cu:171  `boxes.type()` -> `boxes.scalar_type()`
    (AT_DISPATCH_FLOATING_TYPES no longer accepts DeprecatedTypeProperties)
  * nms_kernel.cu:12   `#define N_OFFSETS 36` -> 72 for the OpenLane-V shape
    (the macro is unconditional, -D cannot override it)

Outputs: oracle/_ref/phnet_ref_nms_36*.so and oracle/_ref/phnet_ref_nms_72*.so,
each a pybind module exporting `nms_forward(boxes, scores, thresh, top_k)`
exactly like the reference's `nms_impl` (libs/ops/csrc/nms.cpp:44-61).

Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may load these.
"""
import os
import re
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CSRC = "/root/reference/libs/ops/csrc"
OUT = os.path.join(HERE, "_ref")


def ref_available() -> bool:
    return os.path.isfile(os.path.join(REF_CSRC, "nms_kernel.cu"))


def built(n_off: int) -> str | None:
    if not os.path.isdir(OUT):
        return None
    for f in sorted(os.listdir(OUT)):
        if f.startswith(f"phnet_ref_nms_{n_off}") and f.endswith(".so"):
            return os.path.join(OUT, f)
    return None


def build_one(n_off: int, verbose: bool = False) -> str:
    from torch.utils.cpp_extension import load

    name = f"phnet_ref_nms_{n_off}"
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix=f"phnet_ref_{n_off}_")
    try:
        with open(os.path.join(REF_CSRC, "nms_kernel.cu")) as f:
            cu = f.read()
        cu, n1 = re.subn(r"AT_DISPATCH_FLOATING_TYPES\(boxes\.type\(\)",
                         "AT_DISPATCH_FLOATING_TYPES(boxes.scalar_type()", cu)
        cu, n2 = re.subn(r"#define N_OFFSETS 36\b", f"#define N_OFFSETS {n_off}", cu)
        assert n1 == 1 and n2 == 1, "reference source changed; recipe needs review"
        with open(os.path.join(tmp, "nms_kernel.cu"), "w") as f:
            f.write(cu)
        shutil.copy(os.path.join(REF_CSRC, "nms.cpp"), os.path.join(tmp, "nms.cpp"))
        build_dir = os.path.join(tmp, "build")
        os.makedirs(build_dir)
        os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
        os.environ.setdefault("MAX_JOBS", "4")
        load(
            name=name,
            sources=[os.path.join(tmp, "nms.cpp"), os.path.join(tmp, "nms_kernel.cu")],
            extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo"],
            build_directory=build_dir,
            is_python_module=False,
            verbose=verbose,
        )
        so = os.path.join(build_dir, name + ".so")
        dst = os.path.join(OUT, name + ".so")
        shutil.copy(so, dst)
        return dst
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def build_all(verbose: bool = False) -> dict:
    out = {}
    if not ref_available():
        return out
    for n_off in (36, 72):
        out[n_off] = built(n_off) or build_one(n_off, verbose)
    return out


if __name__ == "__main__":
    print(build_all(verbose="-v" in sys.argv))

// nms_impl.cpp -- the reference's native op surface on top of libphnms.so.
//
// PHNet's libs/ops is a pybind11 module `nms_impl` with ONE function (libs/ops/csrc/nms.cpp:44-61):
//     std::vector<at::Tensor> nms_forward(at::Tensor boxes, at::Tensor scores, float thresh, unsigned long top_k)
// called by libs/ops/nms.py:32-33.  This file builds a module of the same name with the same function, the same argument
// checks (CHECK_CUDA / CHECK_CONTIGUOUS, nms.cpp:40-42,53-54; row width, nms_kernel.cu:154; MAX_COL_BLOCKS, :158; float and
// double only, :171) and the same return value -- three int64 CUDA tensors (keep[N], num_to_keep[], parent_object_index[N])
// -- but forwards to the B200 C ABI (include/phnms.h) instead of launching the reference kernels.  It is a thin adapter:
// ATen for the output allocation, the current stream and the device guard, nothing else.  Per-call host cost is a few
// microseconds (PHNet calls the op once per frame), against ~20 us through ctypes.
#include <torch/extension.h>

#include <c10/cuda/CUDACachingAllocator.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include <mutex>
#include <unordered_map>
#include <vector>

#include "../../include/phnms.h"

namespace {

// small reusable workspace per (device, stream): launches on one stream are ordered, so consecutive calls may share it
at::Tensor &small_workspace(int device, cudaStream_t stream, size_t bytes, const at::TensorOptions &opts) {
    static std::mutex mu;
    static std::unordered_map<unsigned long long, at::Tensor> cache;
    const unsigned long long key = ((unsigned long long)(uintptr_t)stream << 8) ^ (unsigned long long)device;
    std::lock_guard<std::mutex> g(mu);
    at::Tensor &t = cache[key];
    if (!t.defined() || (size_t)t.numel() < bytes) t = at::empty({(long long)std::max<size_t>(bytes, 1 << 16)}, opts.dtype(at::kByte));
    return t;
}

std::vector<at::Tensor> forward_batched(const at::Tensor &boxes, const at::Tensor &scores, const c10::optional<at::Tensor> &n_valid,
                                        double thresh, int64_t top_k, int64_t sort_model, bool single) {
    TORCH_CHECK(boxes.is_cuda(), "boxes must be a CUDA tensor");                     // CHECK_CUDA   (nms.cpp:40)
    TORCH_CHECK(scores.is_cuda(), "scores must be a CUDA tensor");
    TORCH_CHECK(boxes.is_contiguous(), "boxes must be contiguous");                   // CHECK_CONTIGUOUS (nms.cpp:41)
    TORCH_CHECK(boxes.scalar_type() == at::kFloat, "nms_impl (B200): this entry point takes float32 boxes");
    TORCH_CHECK(scores.scalar_type() == at::kFloat && scores.is_contiguous(), "scores must be contiguous float32");
    // (messages are built from strings only: operator<< of an integer crashes inside extensions in some torch wheels'
    // libstdc++ set-ups, reproduced here with a two-line extension)
    TORCH_CHECK(boxes.dim() == (single ? 2 : 3), single ? "boxes must have 2 dimensions" : "boxes must have 3 dimensions");
    TORCH_CHECK(top_k >= 0, "top_k must be non-negative");
    const int64_t F = single ? 1 : boxes.size(0), N = boxes.size(single ? 0 : 1), P = boxes.size(single ? 1 : 2);
    TORCH_CHECK(P >= 6, "Wrong number of offsets. Rows are 5 + n_offsets wide");    // nms_kernel.cu:154
    TORCH_CHECK(scores.numel() == F * N, "scores must have one entry per proposal");
    TORCH_CHECK(scores.device() == boxes.device(), "boxes and scores must be on the same device");
    const c10::cuda::CUDAGuard guard(boxes.device());
    const cudaStream_t stream = c10::cuda::getCurrentCUDAStream(boxes.device().index()).stream();
    const auto lopts = boxes.options().dtype(at::kLong);
    at::Tensor out = at::empty({F * (2 * N + 1)}, lopts);   // one allocation, three views
    // (views built on the storage directly: three dispatcher round trips -- narrow / select -- are ~1.5 us of a ~8 us call; the
    // outputs are int64 and never differentiable, so no view tracking is lost)
    auto view_of = [&out](int64_t offset, c10::IntArrayRef sizes) {
        at::Tensor t = at::detail::make_tensor<c10::TensorImpl>(c10::Storage(out.storage()), out.key_set(), out.dtype());
        t.unsafeGetTensorImpl()->set_storage_offset(offset);
        t.unsafeGetTensorImpl()->set_sizes_contiguous(sizes);
        return t;
    };
    at::Tensor keep = single ? view_of(0, {N}) : view_of(0, {F, N});
    at::Tensor parent = single ? view_of(N, {N}) : view_of(F * N, {F, N});
    at::Tensor num = single ? view_of(2 * N, {}) : view_of(2 * F * N, {F});
    const int32_t *nv = nullptr;
    if (n_valid.has_value()) {
        TORCH_CHECK(!single && n_valid->is_cuda() && n_valid->scalar_type() == at::kInt && n_valid->is_contiguous() && n_valid->numel() == F,
                    "n_valid must be a contiguous int32 CUDA tensor of shape [F]");
        nv = n_valid->data_ptr<int32_t>();
    }
    if (N > 0 && F > 0) {
        const size_t ws_bytes = phnms_workspace_bytes(F, N, (int)(P - 5), nullptr);
        at::Tensor ws;
        void *ws_ptr = nullptr;
        if (ws_bytes > (1u << 20)) {          // large batches: a fresh block, handed back to the allocator in stream order
            ws = at::empty({(long long)ws_bytes}, boxes.options().dtype(at::kByte));
            c10::cuda::CUDACachingAllocator::recordStream(ws.storage().data_ptr(), c10::cuda::getCurrentCUDAStream(boxes.device().index()));
            ws_ptr = ws.data_ptr();
        } else if (ws_bytes > 0) {
            ws_ptr = small_workspace(boxes.device().index(), stream, ws_bytes, boxes.options()).data_ptr();
        }
        const int rc = phnms_forward_f32(boxes.data_ptr<float>(), scores.data_ptr<float>(), nv, F, N, (int)(P - 5), (float)thresh, top_k,
                                         (int)sort_model, keep.data_ptr<int64_t>(), num.data_ptr<int64_t>(), parent.data_ptr<int64_t>(),
                                         ws_ptr, ws_bytes, nullptr, stream);
        TORCH_CHECK(rc == PHNMS_OK, std::string("phnms: ") + phnms_error_string(rc) + " (code " + std::to_string(rc) + ")");
    } else if (F > 0) {
        num.zero_();
    }
    return {keep, num, parent};       // single: keep[N], num_to_keep[] (0-dim), parent_object_index[N]
}

// libs/ops/csrc/nms.cpp:44-48
std::vector<at::Tensor> nms_forward(at::Tensor boxes, at::Tensor scores, float thresh, unsigned long top_k) {
    return forward_batched(boxes, scores, c10::nullopt, (double)thresh, (int64_t)top_k, PHNMS_SORT_TORCH_CUDA, true);
}

// The same call for phnet_b200.ops.nms: None instead of an exception when the arguments are not the plain float32 case, so that
// the Python wrapper needs no checks of its own on the fast path (each attribute lookup there is ~0.1 us of a ~7 us call) and
// sends everything else -- double boxes, strided scores, CPU tensors (which must raise the reference's errors) -- down its
// general path.
pybind11::object nms_forward_or_none(const at::Tensor &boxes, const at::Tensor &scores, double thresh, int64_t top_k) {
    if (!boxes.is_cuda() || !scores.is_cuda() || boxes.scalar_type() != at::kFloat || scores.scalar_type() != at::kFloat ||
        boxes.dim() != 2 || scores.dim() != 1 || !boxes.is_contiguous() || !scores.is_contiguous() || top_k < 0 || boxes.size(1) < 6 ||
        scores.size(0) != boxes.size(0) || scores.device() != boxes.device())
        return pybind11::none();
    return pybind11::cast(forward_batched(boxes, scores, c10::nullopt, thresh, top_k, PHNMS_SORT_TORCH_CUDA, true));
}

std::vector<at::Tensor> nms_forward_batched(at::Tensor boxes, at::Tensor scores, c10::optional<at::Tensor> n_valid, double thresh,
                                            int64_t top_k, int64_t sort_model) {
    return forward_batched(boxes, scores, n_valid, thresh, top_k, sort_model, false);
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("nms_forward", &nms_forward, "nms_forward");   // the reference's export (nms.cpp:59-61)
    m.def("nms_forward_batched", &nms_forward_batched, "F independent nms_forward calls in one launch",
          pybind11::arg("boxes"), pybind11::arg("scores"), pybind11::arg("n_valid") = pybind11::none(), pybind11::arg("thresh") = 50.0,
          pybind11::arg("top_k") = 4, pybind11::arg("sort_model") = 0);
    m.def("nms_forward_or_none", &nms_forward_or_none, "nms_forward, or None when the arguments are not the plain float32 CUDA case");
    m.def("abi_version", []() { return phnms_abi_version(); });
}

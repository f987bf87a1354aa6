// PyTorch C++ extension over the C-ABI (BASELINE.json north_star: "Host code stays Python and calls a PyTorch C++ extension
// through a thin C-ABI"). Registers the capture path's entry points as torch.library operators (namespace `mvmc`) whose CUDA
// implementations do nothing but check tensors, take the current CUDA stream and call libmvmc.so's extern "C" functions
// (include/mvmc.h) with raw device pointers. No arithmetic lives here. Built by multiview_motion_capture_b200/build.py into
// lib/libmvmc_torch.so, loaded with torch.ops.load_library; the ctypes binding in _lib.py stays for the entry points that
// take host pointers.
#include <ATen/ATen.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include <tuple>

#include "mvmc.h"

namespace {

void* cur_stream(const at::Tensor& t) { return (void*)c10::cuda::getCurrentCUDAStream(t.get_device()).stream(); }

void need(const at::Tensor& t, at::ScalarType dt, const char* name) {
    TORCH_CHECK(t.is_cuda(), name, ": CUDA tensor expected (the capture path has no CPU fallback)");
    TORCH_CHECK(t.scalar_type() == dt, name, ": wrong dtype");
    TORCH_CHECK(t.is_contiguous(), name, ": must be contiguous");
}
void ok(int rc, const char* what) {
    TORCH_CHECK(rc == MVMC_OK, what, ": ", mvmc_error_string(rc), rc == MVMC_ERR_CUDA ? mvmc_last_cuda_error() : "");
}

// inverse_kinematics.py:176-199
at::Tensor fk(const at::Tensor& params) {
    need(params, at::kDouble, "params");
    TORCH_CHECK(params.dim() == 2 && params.size(1) == MVMC_N_PARAM, "params [M,68]");
    auto out = at::empty({params.size(0), MVMC_N_B18, 3}, params.options());
    ok(mvmc_fk(params.data_ptr<double>(), (int)params.size(0), out.data_ptr<double>(), cur_stream(params)), "mvmc_fk");
    return out;
}

// mv_math_util.py:152-240
at::Tensor triangulate(const at::Tensor& obs, const at::Tensor& P, const at::Tensor& n_views, double min_score, int64_t refine_nfev) {
    need(obs, at::kDouble, "obs");
    need(P, at::kDouble, "P");
    need(n_views, at::kInt, "n_views");
    TORCH_CHECK(obs.dim() == 4 && obs.size(3) == 3 && P.dim() == 4 && P.size(0) == obs.size(0) && P.size(1) == obs.size(1), "obs [M,V,K,3], P [M,V,3,4]");
    auto out = at::zeros({obs.size(0), obs.size(2), 4}, obs.options());
    ok(mvmc_triangulate(obs.data_ptr<double>(), P.data_ptr<double>(), n_views.data_ptr<int>(), (int)obs.size(0), (int)obs.size(1),
                        (int)obs.size(2), min_score, (int)refine_nfev, out.data_ptr<double>(), cur_stream(obs)), "mvmc_triangulate");
    return out;
}

// inverse_kinematics.py:202-277, 339-433 (PoseSolver.solve, batched)
std::tuple<at::Tensor, at::Tensor, at::Tensor, at::Tensor> ik_solve(const at::Tensor& kps2d, const at::Tensor& Psel, const at::Tensor& n_views,
                                                                    const at::Tensor& x0, const c10::optional<at::Tensor>& birth,
                                                                    const at::Tensor& max_nfev, const c10::optional<at::Tensor>& free_mask) {
    need(kps2d, at::kDouble, "kps2d");
    need(Psel, at::kDouble, "Psel");
    need(n_views, at::kInt, "n_views");
    need(x0, at::kDouble, "x0");
    need(max_nfev, at::kInt, "max_nfev");
    if (birth) need(*birth, at::kByte, "birth");
    if (free_mask) need(*free_mask, at::kByte, "free_mask");
    const int M = (int)kps2d.size(0), V = (int)kps2d.size(1);
    auto ws = at::empty({(int64_t)mvmc_ik_workspace_bytes(M, V) / 8 + 1}, x0.options());
    auto x_out = at::zeros({M, MVMC_N_PARAM}, x0.options());
    auto joints = at::zeros({M, MVMC_N_B18, 3}, x0.options());
    auto info = at::zeros({M, 2, 4}, n_views.options());
    auto cost = at::zeros({M, 2}, x0.options());
    ok(mvmc_ik_solve(kps2d.data_ptr<double>(), Psel.data_ptr<double>(), n_views.data_ptr<int>(), x0.data_ptr<double>(),
                     birth ? birth->data_ptr<uint8_t>() : nullptr, max_nfev.data_ptr<int>(),
                     free_mask ? free_mask->data_ptr<uint8_t>() : nullptr, M, V, ws.data_ptr<double>(), x_out.data_ptr<double>(),
                     joints.data_ptr<double>(), info.data_ptr<int>(), cost.data_ptr<double>(), cur_stream(kps2d)), "mvmc_ik_solve");
    return {x_out, joints, info, cost};
}

// MvTracker.update_4d for a batch of clips (motion_capture.py:873-963): one frame of every clip of the handle
void clips_step(int64_t handle, const at::Tensor& kps, const at::Tensor& n_pose, int64_t frame_idx) {
    need(kps, at::kDouble, "kps");
    need(n_pose, at::kInt, "n_pose");
    ok(mvmc_clips_step(reinterpret_cast<mvmc_clips*>(handle), kps.data_ptr<double>(), n_pose.data_ptr<int>(), (int)frame_idx, cur_stream(kps)),
       "mvmc_clips_step");
}

// compact track records of the last step (the payload of the multi-GPU gather)
void clips_pack_records(int64_t handle, int64_t cap, int64_t clip0, at::Tensor rec, at::Tensor count) {
    need(rec, at::kDouble, "rec");
    need(count, at::kInt, "count");
    ok(mvmc_clips_pack_records(reinterpret_cast<mvmc_clips*>(handle), (int)cap, (int)clip0, rec.data_ptr<double>(), count.data_ptr<int>(), cur_stream(rec)),
       "mvmc_clips_pack_records");
}

}  // namespace

TORCH_LIBRARY(mvmc, m) {
    m.def("fk(Tensor params) -> Tensor");
    m.def("triangulate(Tensor obs, Tensor P, Tensor n_views, float min_score, int refine_nfev) -> Tensor");
    m.def("ik_solve(Tensor kps2d, Tensor Psel, Tensor n_views, Tensor x0, Tensor? birth, Tensor max_nfev, Tensor? free_mask) -> (Tensor, Tensor, Tensor, Tensor)");
    m.def("clips_step(int handle, Tensor kps, Tensor n_pose, int frame_idx) -> ()");
    m.def("clips_pack_records(int handle, int cap, int clip0, Tensor(a!) rec, Tensor(b!) count) -> ()");
}

TORCH_LIBRARY_IMPL(mvmc, CUDA, m) {
    m.impl("fk", &fk);
    m.impl("triangulate", &triangulate);
    m.impl("ik_solve", &ik_solve);
    m.impl("clips_step", &clips_step);
    m.impl("clips_pack_records", &clips_pack_records);
}

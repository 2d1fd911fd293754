// torch_shim.cpp — PyTorch dispatcher glue over the C ABI in include/vision_b200.h.
//
// * defines the ops under our own namespace `vision_b200::` with the reference's schemas
//   (csrc/ops/nms.cpp:27, roi_align.cpp:74-75, roi_pool.cpp:67-68, ps_roi_align.cpp:74-75,
//   deform_conv2d.cpp:101-102) plus `batched_nms` and `resize`, which are Python-only in the
//   reference (torchvision/ops/boxes.py:57-126, transforms/v2/functional/_geometry.py:283-362);
// * `vision_b200::_install(True)` registers the same functions for (`torchvision::<op>`, CUDA)
//   at run time (a heap torch::Library, the dynamic twin of the reference's
//   TORCH_LIBRARY_IMPL(torchvision, CUDA, m) blocks, e.g. cuda/roi_align_kernel.cu:470-477);
//   `_install(False)` destroys it, which re-activates the reference kernels (A/B in-process).
// Tensors are plumbing only: device memory, current stream, allocator.  All checks and error
// strings follow the reference's host functions so its tests read the same.
#include <ATen/ATen.h>
#include <ATen/Context.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include <algorithm>
#include <atomic>
#include <memory>
#include <mutex>
#include <vector>

#include "../../include/vision_b200.h"

namespace {

std::atomic<int> g_nms_semantics{VB200_NMS_CUDA};

int dtype_code(at::ScalarType t, const char* op) {
  switch (t) {
    case at::kFloat: return VB200_F32;
    case at::kHalf: return VB200_F16;
    case at::kBFloat16: return VB200_BF16;
    case at::kDouble: return VB200_F64;
    case at::kByte: return VB200_U8;
    default: TORCH_CHECK(false, op, ": unsupported dtype ", t);
  }
  return -1;
}

void check_rc(int rc, const char* op) {
  TORCH_CHECK(rc == 0, op, ": ", vb200_last_error(), " [vision_b200 rc=", rc, "]");
}

vb200_stream cur_stream() { return (vb200_stream)at::cuda::getCurrentCUDAStream().stream(); }

at::Tensor workspace(size_t bytes, const at::Tensor& like) {
  return at::empty({(int64_t)(bytes ? bytes : 1)}, like.options().dtype(at::kByte));
}

// ---- roi ops -------------------------------------------------------------
void check_roi_inputs(const at::Tensor& input, const at::Tensor& rois) {
  TORCH_CHECK(input.is_cuda(), "input must be a CUDA tensor");
  TORCH_CHECK(rois.is_cuda(), "rois must be a CUDA tensor");
  TORCH_CHECK(rois.dim() == 2 && rois.size(1) == 5, "rois must have shape as Tensor[K, 5]");
  TORCH_CHECK(input.dim() == 4, "input must be a 4-d tensor [N, C, H, W]");
  TORCH_CHECK(input.get_device() == rois.get_device(), "input and rois must be on the same GPU");
  TORCH_CHECK(input.scalar_type() == rois.scalar_type(), "Expected tensor for argument #1 'input' to have the same type as tensor for argument #2 'rois'");
}

at::Tensor roi_align(const at::Tensor& input, const at::Tensor& rois, double spatial_scale, int64_t pooled_height,
                     int64_t pooled_width, int64_t sampling_ratio, bool aligned) {
  check_roi_inputs(input, rois);
  at::cuda::CUDAGuard guard(input.device());
  const int64_t K = rois.size(0), N = input.size(0), C = input.size(1), H = input.size(2), W = input.size(3);
  at::Tensor out = at::empty({K, C, pooled_height, pooled_width}, input.options());
  if (out.numel() == 0) return out;
  const int dt = dtype_code(input.scalar_type(), "roi_align");
  at::Tensor in_c = input.contiguous(), rois_c = rois.contiguous();
  const size_t wsb = vb200_roi_align_workspace_bytes(dt, (int)N, (int)C, (int)H, (int)W, (int)K, (int)pooled_height,
                                                     (int)pooled_width, (int)sampling_ratio);
  at::Tensor ws = workspace(wsb, input);
  check_rc(vb200_roi_align_forward(in_c.data_ptr(), rois_c.data_ptr(), out.data_ptr(), dt, (int)N, (int)C, (int)H, (int)W,
                                   (int)K, (int)pooled_height, (int)pooled_width, spatial_scale, (int)sampling_ratio,
                                   aligned ? 1 : 0, wsb ? ws.data_ptr() : nullptr, wsb, cur_stream()),
           "roi_align");
  return out;
}

// roi_align + all-gather by peer stores: dst_ptrs[0] = this rank's slot of its own gathered buffer, dst_ptrs[1..] = the same
// slot of the peers' buffers; mc_ptr != 0: one NVSwitch multicast address of the slot instead (device pointers as integers).
void roi_align_gather(const at::Tensor& input, const at::Tensor& rois, at::IntArrayRef dst_ptrs, int64_t mc_ptr, double spatial_scale,
                      int64_t pooled_height, int64_t pooled_width, int64_t sampling_ratio, bool aligned) {
  check_roi_inputs(input, rois);
  TORCH_CHECK(dst_ptrs.size() >= 1 && dst_ptrs.size() <= 8, "roi_align_gather: 1..8 destinations");
  at::cuda::CUDAGuard guard(input.device());
  const int64_t K = rois.size(0), N = input.size(0), C = input.size(1), H = input.size(2), W = input.size(3);
  if (K * C * pooled_height * pooled_width == 0) return;
  const int dt = dtype_code(input.scalar_type(), "roi_align");
  at::Tensor in_c = input.contiguous(), rois_c = rois.contiguous();
  const size_t wsb = vb200_roi_align_workspace_bytes(dt, (int)N, (int)C, (int)H, (int)W, (int)K, (int)pooled_height,
                                                     (int)pooled_width, (int)sampling_ratio);
  at::Tensor ws = workspace(wsb, input);
  void* outs[8];
  for (size_t d = 0; d < dst_ptrs.size(); ++d) outs[d] = reinterpret_cast<void*>(static_cast<uintptr_t>(dst_ptrs[d]));
  check_rc(vb200_roi_align_forward_gather(in_c.data_ptr(), rois_c.data_ptr(), outs, (int)dst_ptrs.size(),
                                          reinterpret_cast<void*>(static_cast<uintptr_t>(mc_ptr)), dt, (int)N, (int)C, (int)H, (int)W, (int)K,
                                          (int)pooled_height, (int)pooled_width, spatial_scale, (int)sampling_ratio, aligned ? 1 : 0,
                                          wsb ? ws.data_ptr() : nullptr, wsb, cur_stream()),
           "roi_align_gather");
}

std::tuple<at::Tensor, at::Tensor> roi_pool(const at::Tensor& input, const at::Tensor& rois, double spatial_scale,
                                            int64_t pooled_height, int64_t pooled_width) {
  check_roi_inputs(input, rois);
  at::cuda::CUDAGuard guard(input.device());
  const int64_t K = rois.size(0), N = input.size(0), C = input.size(1), H = input.size(2), W = input.size(3);
  at::Tensor out = at::empty({K, C, pooled_height, pooled_width}, input.options());
  at::Tensor argmax = at::empty({K, C, pooled_height, pooled_width}, input.options().dtype(at::kInt));
  if (out.numel() == 0) return std::make_tuple(out, argmax);
  const int dt = dtype_code(input.scalar_type(), "roi_pool");
  at::Tensor in_c = input.contiguous(), rois_c = rois.contiguous();
  check_rc(vb200_roi_pool_forward(in_c.data_ptr(), rois_c.data_ptr(), out.data_ptr(), argmax.data_ptr<int32_t>(), dt,
                                  (int)N, (int)C, (int)H, (int)W, (int)K, (int)pooled_height, (int)pooled_width,
                                  spatial_scale, cur_stream()),
           "roi_pool");
  return std::make_tuple(out, argmax);
}

std::tuple<at::Tensor, at::Tensor> ps_roi_align(const at::Tensor& input, const at::Tensor& rois, double spatial_scale,
                                                int64_t pooled_height, int64_t pooled_width, int64_t sampling_ratio) {
  check_roi_inputs(input, rois);
  at::cuda::CUDAGuard guard(input.device());
  const int64_t K = rois.size(0), N = input.size(0), C = input.size(1), H = input.size(2), W = input.size(3);
  TORCH_CHECK(C % (pooled_height * pooled_width) == 0,
              "input channels must be a multiple of pooling height * pooling width");
  const int64_t Cout = C / (pooled_height * pooled_width);
  at::Tensor out = at::empty({K, Cout, pooled_height, pooled_width}, input.options());
  at::Tensor mapping = at::empty({K, Cout, pooled_height, pooled_width}, input.options().dtype(at::kInt));
  if (out.numel() == 0) return std::make_tuple(out, mapping);
  const int dt = dtype_code(input.scalar_type(), "ps_roi_align");
  at::Tensor in_c = input.contiguous(), rois_c = rois.contiguous();
  check_rc(vb200_ps_roi_align_forward(in_c.data_ptr(), rois_c.data_ptr(), out.data_ptr(), mapping.data_ptr<int32_t>(), dt,
                                      (int)N, (int)C, (int)H, (int)W, (int)K, (int)pooled_height, (int)pooled_width,
                                      spatial_scale, (int)sampling_ratio, cur_stream()),
           "ps_roi_align");
  return std::make_tuple(out, mapping);
}

std::tuple<at::Tensor, at::Tensor> ps_roi_pool(const at::Tensor& input, const at::Tensor& rois, double spatial_scale,
                                               int64_t pooled_height, int64_t pooled_width) {
  check_roi_inputs(input, rois);
  at::cuda::CUDAGuard guard(input.device());
  const int64_t K = rois.size(0), N = input.size(0), C = input.size(1), H = input.size(2), W = input.size(3);
  TORCH_CHECK(C % (pooled_height * pooled_width) == 0, "input channels must be a multiple of pooling height * pooling width");
  const int64_t Cout = C / (pooled_height * pooled_width);
  at::Tensor out = at::empty({K, Cout, pooled_height, pooled_width}, input.options());
  at::Tensor mapping = at::empty({K, Cout, pooled_height, pooled_width}, input.options().dtype(at::kInt));
  if (out.numel() == 0) return std::make_tuple(out, mapping);
  at::Tensor in_c = input.contiguous(), rois_c = rois.contiguous();
  check_rc(vb200_ps_roi_pool_forward(in_c.data_ptr(), rois_c.data_ptr(), out.data_ptr(), mapping.data_ptr<int32_t>(),
                                     dtype_code(input.scalar_type(), "ps_roi_pool"), (int)N, (int)C, (int)H, (int)W, (int)K,
                                     (int)pooled_height, (int)pooled_width, spatial_scale, cur_stream()),
           "ps_roi_pool");
  return std::make_tuple(out, mapping);
}

at::Tensor ps_roi_pool_backward(const at::Tensor& grad, const at::Tensor& rois, const at::Tensor& channel_mapping, double spatial_scale,
                                int64_t pooled_height, int64_t pooled_width, int64_t batch_size, int64_t channels, int64_t height,
                                int64_t width) {
  TORCH_CHECK(grad.is_cuda() && rois.is_cuda() && channel_mapping.is_cuda(), "grad, rois and channel_mapping must be CUDA tensors");
  TORCH_CHECK(grad.scalar_type() == rois.scalar_type(), "ps_roi_pool_backward: expected grad and rois to have the same dtype");
  at::cuda::CUDAGuard guard(grad.device());
  at::Tensor grad_input = at::empty({batch_size, channels, height, width}, grad.options());
  if (grad_input.numel() == 0) return grad_input;
  at::Tensor g = grad.contiguous(), r = rois.contiguous();
  check_rc(vb200_ps_roi_pool_backward(g.data_ptr(), r.data_ptr(), grad_input.data_ptr(), dtype_code(grad.scalar_type(), "ps_roi_pool_backward"),
                                      (int)batch_size, (int)channels, (int)height, (int)width, (int)r.size(0), (int)pooled_height,
                                      (int)pooled_width, spatial_scale, cur_stream()),
           "ps_roi_pool_backward");
  return grad_input;
}

// ---- fused MultiScaleRoIAlign (torchvision/ops/poolers.py:147-228) ----------------------------------------------
std::tuple<at::Tensor, at::Tensor> multiscale_roi_align(at::TensorList features, const at::Tensor& rois, at::ArrayRef<double> scales,
                                                        int64_t pooled_height, int64_t pooled_width, int64_t sampling_ratio,
                                                        int64_t k_min, int64_t k_max, double canonical_scale, double canonical_level,
                                                        double eps) {
  const int64_t nl = (int64_t)features.size();
  TORCH_CHECK(nl >= 1 && nl <= 8 && (int64_t)scales.size() == nl, "multiscale_roi_align: 1..8 levels with one scale each");
  TORCH_CHECK(rois.is_cuda() && rois.dim() == 2 && rois.size(1) == 5, "rois must have shape as Tensor[K, 5]");
  at::cuda::CUDAGuard guard(rois.device());
  const at::Tensor& f0 = features[0];
  const int64_t B = f0.size(0), C = f0.size(1), K = rois.size(0);
  std::vector<at::Tensor> keep;
  std::vector<const void*> ptrs;
  std::vector<int> hs, ws_;
  for (const at::Tensor& f : features) {
    TORCH_CHECK(f.is_cuda() && f.dim() == 4 && f.size(0) == B && f.size(1) == C && f.scalar_type() == f0.scalar_type() &&
                    f.get_device() == rois.get_device(), "multiscale_roi_align: levels must share device, dtype, batch and channels");
    keep.push_back(f.contiguous());
    ptrs.push_back(keep.back().data_ptr());
    hs.push_back((int)f.size(2));
    ws_.push_back((int)f.size(3));
  }
  TORCH_CHECK(f0.scalar_type() == rois.scalar_type(), "Expected tensor for argument #1 'input' to have the same type as tensor for argument #2 'rois'");
  const int dt = dtype_code(f0.scalar_type(), "multiscale_roi_align");
  TORCH_CHECK(vb200_multiscale_roi_align_supported(dt, (int)nl, hs.data(), ws_.data(), (int)pooled_height, (int)pooled_width,
                                                   (int)sampling_ratio),
              "multiscale_roi_align: unsupported configuration (use the per-level path)");
  at::Tensor out = at::empty({K, C, pooled_height, pooled_width}, f0.options());
  at::Tensor levels = at::empty({K}, f0.options().dtype(at::kInt));
  if (out.numel() == 0) return std::make_tuple(out, levels);
  at::Tensor r = rois.contiguous();
  const size_t wsb = vb200_multiscale_roi_align_workspace_bytes((int)K, (int)nl);
  at::Tensor ws = workspace(wsb, f0);
  std::vector<double> sc(scales.begin(), scales.end());
  check_rc(vb200_multiscale_roi_align_forward(ptrs.data(), hs.data(), ws_.data(), sc.data(), (int)nl, r.data_ptr(), out.data_ptr(),
                                              levels.data_ptr<int32_t>(), dt, (int)B, (int)C, (int)K, (int)pooled_height,
                                              (int)pooled_width, (int)sampling_ratio, (int)k_min, (int)k_max, canonical_scale,
                                              canonical_level, eps, ws.data_ptr(), wsb, cur_stream()),
           "multiscale_roi_align");
  return std::make_tuple(out, levels);
}

// ---- backward of the RoI ops (schemas: roi_align.cpp:76-77, roi_pool.cpp:69-70, ps_roi_align.cpp:76-77) ---------
void check_bwd_inputs(const at::Tensor& grad, const at::Tensor& rois, const char* op) {
  TORCH_CHECK(grad.is_cuda(), "grad must be a CUDA tensor");
  TORCH_CHECK(rois.is_cuda(), "rois must be a CUDA tensor");
  TORCH_CHECK(grad.get_device() == rois.get_device(), op, ": grad and rois must be on the same GPU");
  TORCH_CHECK(grad.scalar_type() == rois.scalar_type(), op, ": expected grad and rois to have the same dtype");
  TORCH_CHECK(rois.dim() == 2 && rois.size(1) == 5, "rois must have shape as Tensor[K, 5]");
}

at::Tensor roi_align_backward(const at::Tensor& grad, const at::Tensor& rois, double spatial_scale, int64_t pooled_height,
                              int64_t pooled_width, int64_t batch_size, int64_t channels, int64_t height, int64_t width,
                              int64_t sampling_ratio, bool aligned) {
  check_bwd_inputs(grad, rois, "roi_align_backward");
  at::cuda::CUDAGuard guard(grad.device());
  // written in full by the kernel (plane by plane); the fallback path zero-fills itself
  at::Tensor grad_input = at::empty({batch_size, channels, height, width}, grad.options());
  if (grad_input.numel() == 0) return grad_input;
  const int dt = dtype_code(grad.scalar_type(), "roi_align_backward");
  at::Tensor g = grad.contiguous(), r = rois.contiguous();
  const size_t wsb = vb200_roi_backward_workspace_bytes((int)r.size(0), (int)pooled_height, (int)pooled_width, (int)sampling_ratio);
  at::Tensor ws = workspace(wsb, grad);
  check_rc(vb200_roi_align_backward(g.data_ptr(), r.data_ptr(), grad_input.data_ptr(), dt, (int)batch_size, (int)channels,
                                    (int)height, (int)width, (int)r.size(0), (int)pooled_height, (int)pooled_width, spatial_scale,
                                    (int)sampling_ratio, aligned ? 1 : 0, at::globalContext().deterministicAlgorithms() ? 1 : 0,
                                    wsb ? ws.data_ptr() : nullptr, wsb, cur_stream()),
           "roi_align_backward");
  return grad_input;
}

at::Tensor roi_pool_backward(const at::Tensor& grad, const at::Tensor& rois, const at::Tensor& argmax, double spatial_scale,
                             int64_t pooled_height, int64_t pooled_width, int64_t batch_size, int64_t channels, int64_t height,
                             int64_t width) {
  check_bwd_inputs(grad, rois, "roi_pool_backward");
  TORCH_CHECK(argmax.is_cuda() && argmax.scalar_type() == at::kInt, "argmax must be a CUDA int32 tensor");
  at::cuda::CUDAGuard guard(grad.device());
  at::Tensor grad_input = at::empty({batch_size, channels, height, width}, grad.options());
  if (grad_input.numel() == 0) return grad_input;
  const int dt = dtype_code(grad.scalar_type(), "roi_pool_backward");
  at::Tensor g = grad.contiguous(), r = rois.contiguous(), am = argmax.contiguous();
  const size_t wsb = vb200_roi_backward_workspace_bytes((int)r.size(0), (int)pooled_height, (int)pooled_width, 1);
  at::Tensor ws = workspace(wsb, grad);
  check_rc(vb200_roi_pool_backward(g.data_ptr(), r.data_ptr(), am.data_ptr<int32_t>(), grad_input.data_ptr(), dt, (int)batch_size,
                                   (int)channels, (int)height, (int)width, (int)r.size(0), (int)pooled_height, (int)pooled_width,
                                   spatial_scale, at::globalContext().deterministicAlgorithms() ? 1 : 0,
                                   wsb ? ws.data_ptr() : nullptr, wsb, cur_stream()),
           "roi_pool_backward");
  return grad_input;
}

at::Tensor ps_roi_align_backward(const at::Tensor& grad, const at::Tensor& rois, const at::Tensor& channel_mapping,
                                 double spatial_scale, int64_t pooled_height, int64_t pooled_width, int64_t sampling_ratio,
                                 int64_t batch_size, int64_t channels, int64_t height, int64_t width) {
  check_bwd_inputs(grad, rois, "ps_roi_align_backward");
  TORCH_CHECK(channel_mapping.is_cuda(), "channel_mapping must be a CUDA tensor");
  at::cuda::CUDAGuard guard(grad.device());
  at::Tensor grad_input = at::empty({batch_size, channels, height, width}, grad.options());
  if (grad_input.numel() == 0) return grad_input;
  const int dt = dtype_code(grad.scalar_type(), "ps_roi_align_backward");
  at::Tensor g = grad.contiguous(), r = rois.contiguous(), cm = channel_mapping.contiguous();
  const size_t wsb = vb200_roi_backward_workspace_bytes((int)r.size(0), (int)pooled_height, (int)pooled_width, (int)sampling_ratio);
  at::Tensor ws = workspace(wsb, grad);
  check_rc(vb200_ps_roi_align_backward(g.data_ptr(), r.data_ptr(), cm.data_ptr<int32_t>(), grad_input.data_ptr(), dt, (int)batch_size,
                                       (int)channels, (int)height, (int)width, (int)r.size(0), (int)pooled_height,
                                       (int)pooled_width, spatial_scale, (int)sampling_ratio,
                                       at::globalContext().deterministicAlgorithms() ? 1 : 0, wsb ? ws.data_ptr() : nullptr, wsb,
                                       cur_stream()),
           "ps_roi_align_backward");
  return grad_input;
}

// ---- nms -------------------------------------------------------------------
void check_nms_inputs(const at::Tensor& dets, const at::Tensor& scores) {
  TORCH_CHECK(dets.is_cuda(), "dets must be a CUDA tensor");
  TORCH_CHECK(scores.is_cuda(), "scores must be a CUDA tensor");
  TORCH_CHECK(dets.dim() == 2, "boxes should be a 2d tensor, got ", dets.dim(), "D");
  TORCH_CHECK(dets.size(1) == 4, "boxes should have 4 elements in dimension 1, got ", dets.size(1));
  TORCH_CHECK(scores.dim() == 1, "scores should be a 1d tensor, got ", scores.dim(), "D");
  TORCH_CHECK(dets.size(0) == scores.size(0), "boxes and scores should have same number of elements in ",
              "dimension 0, got ", dets.size(0), " and ", scores.size(0));
}

// The reference instantiates float / double / Half (cuda/nms_kernel.cu:32-54); all three run natively, each
// with the arithmetic of its compiled reference kernel.  Other dtypes are rejected as the reference's dispatch does.
// A contiguous view whose storage offset breaks the one-box alignment of the vector loads is copied.
int nms_dtype(const at::Tensor& t, const char* op) {
  const auto st = t.scalar_type();
  TORCH_CHECK(st == at::kFloat || st == at::kDouble || st == at::kHalf, op, ": \"nms_kernel\" not implemented for '", st, "'");
  return st == at::kDouble ? VB200_F64 : st == at::kHalf ? VB200_F16 : VB200_F32;
}
at::Tensor nms_operand(const at::Tensor& t, size_t align) {
  at::Tensor c = t.contiguous();
  if (((uintptr_t)c.data_ptr() % align) != 0) c = c.clone();
  return c;
}

at::Tensor nms(const at::Tensor& dets, const at::Tensor& scores, double iou_threshold) {
  check_nms_inputs(dets, scores);
  TORCH_CHECK(dets.scalar_type() == scores.scalar_type(), "dets should have the same type as scores");
  at::cuda::CUDAGuard guard(dets.device());
  if (dets.numel() == 0) return at::empty({0}, dets.options().dtype(at::kLong));
  const int dt = nms_dtype(dets, "nms");
  at::Tensor boxes = nms_operand(dets, 4 * dets.element_size()), sc = nms_operand(scores, scores.element_size());
  const int64_t n = boxes.size(0);
  const size_t wsb = vb200_nms_workspace_bytes(n);
  at::Tensor ws = workspace(wsb, boxes);
  at::Tensor keep = at::empty({n}, boxes.options().dtype(at::kLong));
  at::Tensor count = at::empty({1}, boxes.options().dtype(at::kLong));
  check_rc(vb200_nms(boxes.data_ptr(), sc.data_ptr(), dt, n, iou_threshold, dt == VB200_F16 ? VB200_NMS_CUDA : g_nms_semantics.load(), ws.data_ptr(),
                     wsb, keep.data_ptr<int64_t>(), count.data_ptr<int64_t>(), cur_stream()),
           "nms");
  const int64_t k = count.item<int64_t>();   // the reference's masked_select sync (nms_kernel.cu:257)
  return keep.narrow(0, 0, k);
}

// Device-side result of batched_nms: keep [n] (first *count entries valid, the rest unspecified) and count [1], both on
// the device - no host synchronisation.  The sharded path gathers these padded lists with one collective.
std::tuple<at::Tensor, at::Tensor> batched_nms_padded(const at::Tensor& dets, const at::Tensor& scores, const at::Tensor& idxs,
                                                      double iou_threshold) {
  check_nms_inputs(dets, scores);
  TORCH_CHECK(idxs.is_cuda(), "idxs must be a CUDA tensor");
  TORCH_CHECK(idxs.dim() == 1 && idxs.size(0) == dets.size(0), "idxs should be a 1d tensor with one entry per box");
  at::cuda::CUDAGuard guard(dets.device());
  const int64_t n = dets.size(0);
  at::Tensor keep = at::empty({n}, dets.options().dtype(at::kLong));
  at::Tensor count = at::zeros({1}, dets.options().dtype(at::kLong));
  if (dets.numel() == 0) return std::make_tuple(keep, count);
  TORCH_CHECK(dets.scalar_type() == scores.scalar_type(), "boxes should have the same type as scores");
  const int dt = nms_dtype(dets, "batched_nms");
  at::Tensor boxes = nms_operand(dets, 4 * dets.element_size()), sc = nms_operand(scores, scores.element_size());
  at::Tensor cls = idxs.to(at::kLong).contiguous();
  const size_t wsb = vb200_batched_nms_workspace_bytes(n);
  at::Tensor ws = workspace(wsb, boxes);
  // class ids are sorted as 16-bit keys (the common case); ids outside [0, 65536) make the kernel report count = -1
  check_rc(vb200_batched_nms(boxes.data_ptr(), sc.data_ptr(), cls.data_ptr<int64_t>(), dt, n, iou_threshold,
                             dt == VB200_F16 ? VB200_NMS_CUDA : g_nms_semantics.load(), VB200_BNMS_AUTO, ws.data_ptr(), wsb,
                             keep.data_ptr<int64_t>(), count.data_ptr<int64_t>(), cur_stream()),
           "batched_nms");
  return std::make_tuple(keep, count);
}

at::Tensor batched_nms(const at::Tensor& dets, const at::Tensor& scores, const at::Tensor& idxs, double iou_threshold) {
  check_nms_inputs(dets, scores);
  TORCH_CHECK(idxs.is_cuda(), "idxs must be a CUDA tensor");
  TORCH_CHECK(idxs.dim() == 1 && idxs.size(0) == dets.size(0), "idxs should be a 1d tensor with one entry per box");
  at::cuda::CUDAGuard guard(dets.device());
  if (dets.numel() == 0) return at::empty({0}, dets.options().dtype(at::kLong));
  TORCH_CHECK(dets.scalar_type() == scores.scalar_type(), "boxes should have the same type as scores");
  const int dt = nms_dtype(dets, "batched_nms");
  at::Tensor boxes = nms_operand(dets, 4 * dets.element_size()), sc = nms_operand(scores, scores.element_size());
  at::Tensor cls = idxs.to(at::kLong).contiguous();
  const int64_t n = boxes.size(0);
  const size_t wsb = vb200_batched_nms_workspace_bytes(n);
  at::Tensor ws = workspace(wsb, boxes);
  at::Tensor keep = at::empty({n}, boxes.options().dtype(at::kLong));
  at::Tensor count = at::empty({1}, boxes.options().dtype(at::kLong));
  int64_t k = -1;
  for (int attempt = 0; attempt < 2 && k < 0; ++attempt) {
    // first attempt speculates 16-bit class ids; -1 asks for the wide-key repeat (arbitrary int64 ids)
    const int strategy = VB200_BNMS_AUTO | (attempt ? VB200_BNMS_WIDE_KEYS : 0);
    check_rc(vb200_batched_nms(boxes.data_ptr(), sc.data_ptr(), cls.data_ptr<int64_t>(), dt, n, iou_threshold,
                               dt == VB200_F16 ? VB200_NMS_CUDA : g_nms_semantics.load(), strategy, ws.data_ptr(), wsb,
                               keep.data_ptr<int64_t>(), count.data_ptr<int64_t>(), cur_stream()),
             "batched_nms");
    k = count.item<int64_t>();
  }
  TORCH_CHECK(k >= 0, "batched_nms: internal error (negative kept count)");
  return keep.narrow(0, 0, k);
}

// ---- detection post-processing around batched_nms (roi_heads.py:700-737, rpn.py:273-298) ---------------------------
std::tuple<at::Tensor, at::Tensor, at::Tensor> detection_postprocess(const at::Tensor& boxes, const at::Tensor& scores,
                                                                     const at::Tensor& labels, double img_h, double img_w,
                                                                     double score_thresh, bool score_inclusive, double min_size,
                                                                     double nms_thresh, int64_t topk) {
  check_nms_inputs(boxes, scores);
  TORCH_CHECK(labels.is_cuda() && labels.dim() == 1 && labels.size(0) == boxes.size(0), "labels should be a 1d tensor with one entry per box");
  TORCH_CHECK(boxes.scalar_type() == at::kFloat && scores.scalar_type() == at::kFloat, "detection_postprocess: float32 boxes and scores");
  at::cuda::CUDAGuard guard(boxes.device());
  const int64_t n = boxes.size(0);
  const int64_t cap = std::min<int64_t>(n, std::max<int64_t>(topk, 0));
  at::Tensor ob = at::empty({cap, 4}, boxes.options()), os = at::empty({cap}, scores.options());
  at::Tensor ol = at::empty({cap}, boxes.options().dtype(at::kLong));
  if (cap == 0) return std::make_tuple(ob, os, ol);
  at::Tensor b = nms_operand(boxes, 16), s = nms_operand(scores, 4), l = labels.to(at::kLong).contiguous();
  const size_t wsb = vb200_detection_postprocess_workspace_bytes(n);
  at::Tensor ws = workspace(wsb, boxes);
  int64_t count = 0;
  check_rc(vb200_detection_postprocess(b.data_ptr(), s.data_ptr(), l.data_ptr<int64_t>(), VB200_F32, n, img_h, img_w, score_thresh,
                                       score_inclusive ? 1 : 0, min_size, nms_thresh, topk, g_nms_semantics.load(), ws.data_ptr(), wsb,
                                       ob.data_ptr(), os.data_ptr(), ol.data_ptr<int64_t>(), &count, cur_stream()),
           "detection_postprocess");
  return std::make_tuple(ob.narrow(0, 0, count), os.narrow(0, 0, count), ol.narrow(0, 0, count));
}

// ---- deform_conv2d ---------------------------------------------------------
// Packed weights are cached per weight tensor: the key is the TensorImpl (held weakly, so a recycled address cannot
// alias) plus its version counter (an in-place update of the parameter invalidates the entry) plus the generation of the
// VB200_* overrides (VB200_DCN_CTA2 / VB200_DCN_BN change the packed layout).
struct PackedWeight {
  c10::weak_intrusive_ptr<c10::TensorImpl> impl;
  uint32_t version;
  int dtype;
  int env_gen;
  at::Tensor packed;
};
std::mutex g_pack_mu;
std::vector<PackedWeight> g_pack_cache;

at::Tensor packed_weight_for(const at::Tensor& weight_c, int dt, int c_in, int c_out, int kh, int kw, int groups, int offset_groups) {
  const size_t bytes = vb200_deform_conv2d_packed_weight_bytes(dt, c_in, c_out, kh, kw, groups, offset_groups);
  if (bytes == 0 || weight_c.is_inference()) return at::Tensor();
  c10::TensorImpl* impl = weight_c.unsafeGetTensorImpl();
  const uint32_t version = (uint32_t)weight_c._version();
  const int env_gen = vb200_env_generation();
  std::lock_guard<std::mutex> lk(g_pack_mu);
  for (size_t i = 0; i < g_pack_cache.size(); ++i) {
    auto locked = g_pack_cache[i].impl.lock();
    if (!locked) { g_pack_cache.erase(g_pack_cache.begin() + i); --i; continue; }      // the weight died
    if (locked.get() == impl && g_pack_cache[i].dtype == dt) {
      if (g_pack_cache[i].version == version && g_pack_cache[i].env_gen == env_gen) return g_pack_cache[i].packed;
      g_pack_cache.erase(g_pack_cache.begin() + i);                                     // updated in place: re-pack
      break;
    }
  }
  at::Tensor packed = at::empty({(int64_t)bytes}, weight_c.options().dtype(at::kByte));
  check_rc(vb200_deform_conv2d_pack_weight(weight_c.data_ptr(), packed.data_ptr(), dt, c_in, c_out, kh, kw, groups, offset_groups,
                                           (vb200_stream)at::cuda::getCurrentCUDAStream().stream()),
           "deform_conv2d");
  if (g_pack_cache.size() >= 32) g_pack_cache.erase(g_pack_cache.begin());
  g_pack_cache.push_back(PackedWeight{c10::weak_intrusive_ptr<c10::TensorImpl>(c10::intrusive_ptr<c10::TensorImpl>::reclaim_copy(impl)),
                                      version, dt, env_gen, packed});
  return packed;
}

// dst_ptrs == nullptr: allocate and return the output.  Otherwise (fused all-gather): write to dst_ptrs[0] (this rank's slot of its
// own gathered buffer) and dst_ptrs[1..] (the same slot of the peers' buffers); returns an undefined tensor.
at::Tensor deform_conv2d_impl(const at::Tensor& input, const at::Tensor& weight, const at::Tensor& offset,
                              const at::Tensor& mask, const at::Tensor& bias, int64_t stride_h, int64_t stride_w,
                              int64_t pad_h, int64_t pad_w, int64_t dilation_h, int64_t dilation_w, int64_t n_weight_grps,
                              int64_t n_offset_grps, bool use_mask, const at::IntArrayRef* dst_ptrs) {
  // a channels-last input is handed to the tensor-core path as it is (no NCHW -> NHWC staging pass)
  const bool nhwc_in = input.dim() == 4 && !input.is_contiguous() && input.is_contiguous(at::MemoryFormat::ChannelsLast);
  at::Tensor input_c = nhwc_in ? input : input.contiguous(), offset_c = offset.contiguous(), weight_c = weight.contiguous();
  at::Tensor mask_c = mask.contiguous(), bias_c = bias.contiguous();
  TORCH_CHECK(input_c.ndimension() == 4);
  TORCH_CHECK(offset_c.ndimension() == 4);
  TORCH_CHECK(!use_mask || mask_c.ndimension() == 4);
  TORCH_CHECK(weight_c.ndimension() == 4);
  TORCH_CHECK(input_c.is_cuda(), "input must be a CUDA tensor");
  at::cuda::CUDAGuard guard(input_c.device());

  const int64_t batch = input_c.size(0), c_in = input_c.size(1), in_h = input_c.size(2), in_w = input_c.size(3);
  const int64_t c_out = weight_c.size(0), kh = weight_c.size(2), kw = weight_c.size(3);
  TORCH_CHECK(kh > 0 && kw > 0, "weight_h: ", kh, " weight_w: ", kw);
  TORCH_CHECK(stride_h > 0 && stride_w > 0, "stride_h: ", stride_h, " stride_w: ", stride_w);
  TORCH_CHECK(pad_h >= 0 && pad_w >= 0, "pad_h: ", pad_h, " pad_w: ", pad_w);
  TORCH_CHECK(dilation_h > 0 && dilation_w > 0, "dilation_h: ", dilation_h, " dilation_w: ", dilation_w);
  const int64_t ker_h = dilation_h * (kh - 1) + 1, ker_w = dilation_w * (kw - 1) + 1;
  const int64_t out_h = ((in_h + 2 * pad_h - ker_h) / stride_h) + 1;
  const int64_t out_w = ((in_w + 2 * pad_w - ker_w) / stride_w) + 1;
  TORCH_CHECK(n_weight_grps > 0 && n_offset_grps > 0);
  TORCH_CHECK(weight_c.size(1) * n_weight_grps == c_in);
  TORCH_CHECK(c_out % n_weight_grps == 0);
  TORCH_CHECK(offset_c.size(1) == n_offset_grps * 2 * kh * kw, "offset.shape[1] is not valid: got: ", offset_c.size(1),
              " expected: ", n_offset_grps * 2 * kh * kw);
  TORCH_CHECK(!use_mask || mask_c.size(1) == n_offset_grps * kh * kw, "mask.shape[1] is not valid: got: ",
              mask_c.size(1), " expected: ", n_offset_grps * kh * kw);
  TORCH_CHECK(c_in % n_offset_grps == 0);
  TORCH_CHECK(offset_c.size(0) == batch, "invalid batch size of offset");
  TORCH_CHECK(offset_c.size(2) == out_h && offset_c.size(3) == out_w, "offset output dims: (", offset_c.size(2), ", ",
              offset_c.size(3), ") - computed output dims: (", out_h, ", ", out_w, ")");
  TORCH_CHECK(mask_c.size(0) == batch, "invalid batch size of mask");
  TORCH_CHECK(!use_mask || (mask_c.size(2) == out_h && mask_c.size(3) == out_w), "mask output dims: (", mask_c.size(2),
              ", ", mask_c.size(3), ") - computed output dims: (", out_h, ", ", out_w, ")");
  TORCH_CHECK(out_h > 0 && out_w > 0, "Calculated output size too small - out_h: ", out_h, " out_w: ", out_w);
  const auto st = input_c.scalar_type();
  TORCH_CHECK(weight_c.scalar_type() == st && offset_c.scalar_type() == st && (!use_mask || mask_c.scalar_type() == st) &&
                  bias_c.scalar_type() == st,
              "deform_conv2d: all tensors must share one dtype");

  at::Tensor out = dst_ptrs ? at::Tensor() : at::empty({batch, c_out, out_h, out_w}, input_c.options());
  if (batch == 0 || batch * c_out * out_h * out_w == 0) return out;
  const int dt = dtype_code(st, "deform_conv2d");
  TORCH_CHECK(dt == VB200_F32 || dt == VB200_F16 || dt == VB200_BF16 || dt == VB200_F64, "deform_conv2d: unsupported dtype ", st);
  TORCH_CHECK(bias_c.numel() == c_out, "bias must have one entry per output channel");
  const size_t wsb = vb200_deform_conv2d_workspace_bytes(dt, (int)batch, (int)c_in, (int)in_h, (int)in_w, (int)c_out, (int)kh,
                                                         (int)kw, (int)out_h, (int)out_w, (int)n_weight_grps,
                                                         (int)n_offset_grps);
  at::Tensor ws = workspace(wsb, input_c);
  at::Tensor packed = wsb ? packed_weight_for(weight_c, dt, (int)c_in, (int)c_out, (int)kh, (int)kw, (int)n_weight_grps, (int)n_offset_grps)
                          : at::Tensor();
  const bool nhwc_ok = nhwc_in && wsb > 0 && ((uintptr_t)input_c.data_ptr() % 16) == 0;      // wsb > 0 <=> tensor-core path
  if (nhwc_in && !nhwc_ok) input_c = input.contiguous();
  void* outs[8];
  int n_outs = 1;
  if (dst_ptrs) {
    TORCH_CHECK(dst_ptrs->size() >= 1 && dst_ptrs->size() <= 8, "deform_conv2d_gather: 1..8 destinations");
    n_outs = (int)dst_ptrs->size();
    for (int d = 0; d < n_outs; ++d) outs[d] = reinterpret_cast<void*>(static_cast<uintptr_t>((*dst_ptrs)[d]));
  } else {
    outs[0] = out.data_ptr();
  }
  check_rc(vb200_deform_conv2d_forward_gather(input_c.data_ptr(), weight_c.data_ptr(), packed.defined() ? packed.data_ptr() : nullptr,
                                              nhwc_ok ? 1 : 0, offset_c.data_ptr(), use_mask ? mask_c.data_ptr() : nullptr,
                                              bias_c.data_ptr(), outs, n_outs, dt, (int)batch, (int)c_in, (int)in_h, (int)in_w, (int)c_out,
                                              (int)kh, (int)kw, (int)stride_h, (int)stride_w, (int)pad_h, (int)pad_w, (int)dilation_h,
                                              (int)dilation_w, (int)n_weight_grps, (int)n_offset_grps, use_mask ? 1 : 0,
                                              wsb ? ws.data_ptr() : nullptr, wsb, cur_stream()),
           "deform_conv2d");
  return out;
}

at::Tensor deform_conv2d(const at::Tensor& input, const at::Tensor& weight, const at::Tensor& offset,
                         const at::Tensor& mask, const at::Tensor& bias, int64_t stride_h, int64_t stride_w,
                         int64_t pad_h, int64_t pad_w, int64_t dilation_h, int64_t dilation_w, int64_t n_weight_grps,
                         int64_t n_offset_grps, bool use_mask) {
  return deform_conv2d_impl(input, weight, offset, mask, bias, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w, n_weight_grps,
                            n_offset_grps, use_mask, nullptr);
}

void deform_conv2d_gather(const at::Tensor& input, const at::Tensor& weight, const at::Tensor& offset, const at::Tensor& mask,
                          const at::Tensor& bias, at::IntArrayRef dst_ptrs, int64_t stride_h, int64_t stride_w, int64_t pad_h, int64_t pad_w,
                          int64_t dilation_h, int64_t dilation_w, int64_t n_weight_grps, int64_t n_offset_grps, bool use_mask) {
  deform_conv2d_impl(input, weight, offset, mask, bias, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w, n_weight_grps,
                     n_offset_grps, use_mask, &dst_ptrs);
}

// ---- deform_conv2d backward (schema csrc/ops/deform_conv2d.cpp:103-104; reference deform_conv2d_kernel.cu:647-1033) -------
// Two plain GEMMs (cuBLAS through at::matmul: weight^T x grad_out -> dcol; grad_out x columns^T -> grad_weight) around two
// kernels of ours: the fused grad_input / grad_offset / grad_mask pass and the column sampler (deform_conv2d_bwd.cu).
std::tuple<at::Tensor, at::Tensor, at::Tensor, at::Tensor, at::Tensor> deform_conv2d_backward(
    const at::Tensor& grad, const at::Tensor& input, const at::Tensor& weight, const at::Tensor& offset, const at::Tensor& mask,
    const at::Tensor& bias, int64_t stride_h, int64_t stride_w, int64_t pad_h, int64_t pad_w, int64_t dilation_h, int64_t dilation_w,
    int64_t n_weight_grps, int64_t n_offset_grps, bool use_mask) {
  TORCH_CHECK(grad.is_cuda() && input.is_cuda(), "deform_conv2d_backward: CUDA tensors expected");
  at::cuda::CUDAGuard guard(input.device());
  at::Tensor grad_c = grad.contiguous(), input_c = input.contiguous(), weight_c = weight.contiguous(), offset_c = offset.contiguous();
  at::Tensor mask_c = mask.contiguous();
  const int64_t B = input_c.size(0), C_in = input_c.size(1), H = input_c.size(2), W = input_c.size(3);
  const int64_t C_out = weight_c.size(0), cin_g = weight_c.size(1), kh = weight_c.size(2), kw = weight_c.size(3), KK = kh * kw;
  TORCH_CHECK(n_weight_grps > 0 && n_offset_grps > 0 && cin_g * n_weight_grps == C_in && C_out % n_weight_grps == 0 && C_in % n_offset_grps == 0,
              "deform_conv2d_backward: channels not divisible by groups");
  const auto st = input_c.scalar_type();
  TORCH_CHECK(grad_c.scalar_type() == st && weight_c.scalar_type() == st && offset_c.scalar_type() == st && (!use_mask || mask_c.scalar_type() == st),
              "deform_conv2d_backward: all tensors must share one dtype");
  const int dt = dtype_code(st, "deform_conv2d_backward");
  TORCH_CHECK(dt == VB200_F32 || dt == VB200_F64 || dt == VB200_F16 || dt == VB200_BF16, "deform_conv2d_backward: unsupported dtype ", st);
  at::Tensor grad_input = at::zeros_like(input_c), grad_offset = at::empty_like(offset_c), grad_weight = at::zeros_like(weight_c);
  at::Tensor grad_mask = use_mask ? at::empty_like(mask_c) : at::zeros_like(mask_c);
  at::Tensor grad_bias = at::ones_like(bias) * (grad_c.numel() ? grad_c.sum({0, 2, 3}) : at::zeros_like(bias));   // deform_conv2d_kernel.cu:1231
  if (B == 0 || grad_c.numel() == 0) return std::make_tuple(grad_input, grad_weight, grad_offset, grad_mask, grad_bias);
  const int64_t out_h = grad_c.size(2), out_w = grad_c.size(3), HWo = out_h * out_w, cout_g = C_out / n_weight_grps;
  TORCH_CHECK(grad_c.size(0) == B && grad_c.size(1) == C_out && offset_c.size(2) == out_h && offset_c.size(3) == out_w,
              "deform_conv2d_backward: grad / offset shapes do not match the forward geometry");
  const int64_t per_img = C_in * KK * HWo * (int64_t)input_c.element_size();
  const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(B, (int64_t)(1ll << 30) / std::max<int64_t>(per_img, 1)));
  const bool low = st == at::kHalf || st == at::kBFloat16;
  at::Tensor gw_acc = low ? at::zeros({C_out, cin_g * KK}, input_c.options().dtype(at::kFloat)) : grad_weight.view({C_out, cin_g * KK});
  at::Tensor w2 = weight_c.view({C_out, cin_g * KK});
  for (int64_t b0 = 0; b0 < B; b0 += chunk) {
    const int64_t nb = std::min(chunk, B - b0);
    at::Tensor g = grad_c.narrow(0, b0, nb).view({nb, C_out, HWo});
    at::Tensor buf = at::empty({nb, C_in * KK, HWo}, input_c.options());
    // dcol = weight^T x grad_out, per weight group
    for (int64_t grp = 0; grp < n_weight_grps; ++grp) {
      at::Tensor wt = w2.narrow(0, grp * cout_g, cout_g).t();                                  // [cin_g*KK, cout_g]
      at::Tensor d = at::matmul(wt, g.narrow(1, grp * cout_g, cout_g));                        // [nb, cin_g*KK, HWo]
      if (n_weight_grps == 1) buf = d; else buf.narrow(1, grp * cin_g * KK, cin_g * KK).copy_(d);
    }
    at::Tensor in_b = input_c.narrow(0, b0, nb), off_b = offset_c.narrow(0, b0, nb);
    at::Tensor gi_b = grad_input.narrow(0, b0, nb), go_b = grad_offset.narrow(0, b0, nb);
    const void* mk = use_mask ? mask_c.narrow(0, b0, nb).data_ptr() : nullptr;
    void* gm = use_mask ? grad_mask.narrow(0, b0, nb).data_ptr() : nullptr;
    check_rc(vb200_deform_conv2d_backward_inputs(buf.data_ptr(), in_b.data_ptr(), off_b.data_ptr(), mk, gi_b.data_ptr(), go_b.data_ptr(), gm,
                                                 dt, (int)nb, (int)C_in, (int)H, (int)W, (int)kh, (int)kw, (int)stride_h, (int)stride_w,
                                                 (int)pad_h, (int)pad_w, (int)dilation_h, (int)dilation_w, (int)n_offset_grps,
                                                 use_mask ? 1 : 0, cur_stream()),
             "deform_conv2d_backward");
    // columns for grad_weight (the buffer is reused)
    check_rc(vb200_deform_conv2d_sample_columns(in_b.data_ptr(), off_b.data_ptr(), mk, buf.data_ptr(), dt, (int)nb, (int)C_in, (int)H, (int)W,
                                                (int)kh, (int)kw, (int)stride_h, (int)stride_w, (int)pad_h, (int)pad_w, (int)dilation_h,
                                                (int)dilation_w, (int)n_offset_grps, use_mask ? 1 : 0, cur_stream()),
             "deform_conv2d_backward");
    for (int64_t grp = 0; grp < n_weight_grps; ++grp) {
      at::Tensor cols = buf.narrow(1, grp * cin_g * KK, cin_g * KK);                           // [nb, cin_g*KK, HWo]
      at::Tensor gw = at::matmul(g.narrow(1, grp * cout_g, cout_g), cols.transpose(1, 2));     // [nb, cout_g, cin_g*KK]
      at::Tensor acc = gw_acc.narrow(0, grp * cout_g, cout_g);
      acc.add_(low ? gw.to(at::kFloat).sum(0) : gw.sum(0));
    }
  }
  if (low) grad_weight.view({C_out, cin_g * KK}).copy_(gw_acc);
  return std::make_tuple(grad_input, grad_weight, grad_offset, grad_mask, grad_bias);
}

// ---- resize ----------------------------------------------------------------
// input [..., H, W] -> [..., out_h, out_w]; mode 0 bilinear / 1 bicubic (align_corners=False).
at::Tensor resize(const at::Tensor& input, int64_t out_h, int64_t out_w, int64_t mode, bool antialias) {
  TORCH_CHECK(input.is_cuda(), "input must be a CUDA tensor");
  TORCH_CHECK(input.dim() >= 2, "resize: input must have at least 2 dimensions");
  TORCH_CHECK(out_h > 0 && out_w > 0, "resize: output size must be positive");
  at::cuda::CUDAGuard guard(input.device());
  at::Tensor in_c = input.contiguous();
  std::vector<int64_t> shape(in_c.sizes().begin(), in_c.sizes().end());
  const int64_t in_h = shape[shape.size() - 2], in_w = shape[shape.size() - 1];
  TORCH_CHECK(in_h > 0 && in_w > 0, "resize: empty spatial dimensions");
  shape[shape.size() - 2] = out_h;
  shape[shape.size() - 1] = out_w;
  at::Tensor out = at::empty(shape, in_c.options());
  if (out.numel() == 0) return out;
  const int64_t planes = in_c.numel() / (in_h * in_w);
  const int dt = dtype_code(in_c.scalar_type(), "resize");
  check_rc(vb200_resize(in_c.data_ptr(), out.data_ptr(), dt, planes, (int)in_h, (int)in_w, (int)out_h, (int)out_w, (int)mode,
                        antialias ? 1 : 0, cur_stream()),
           "resize");
  return out;
}

// resize + all-gather by peer stores: dst_ptrs[0] = this rank's slot of its own gathered buffer, dst_ptrs[1..] = the same slot
// of the peers' buffers (device pointers as integers, e.g. _SymmetricMemory.buffer_ptrs[r] + slot offset).
void resize_gather(const at::Tensor& input, at::IntArrayRef dst_ptrs, int64_t out_h, int64_t out_w, int64_t mode, bool antialias) {
  TORCH_CHECK(input.is_cuda(), "input must be a CUDA tensor");
  TORCH_CHECK(input.dim() >= 2, "resize: input must have at least 2 dimensions");
  TORCH_CHECK(out_h > 0 && out_w > 0, "resize: output size must be positive");
  TORCH_CHECK(dst_ptrs.size() >= 1 && dst_ptrs.size() <= 8, "resize_gather: 1..8 destinations");
  at::cuda::CUDAGuard guard(input.device());
  at::Tensor in_c = input.contiguous();
  const int64_t in_h = in_c.size(-2), in_w = in_c.size(-1);
  TORCH_CHECK(in_h > 0 && in_w > 0, "resize: empty spatial dimensions");
  if (in_c.numel() == 0) return;
  const int64_t planes = in_c.numel() / (in_h * in_w);
  void* outs[8];
  for (size_t d = 0; d < dst_ptrs.size(); ++d) outs[d] = reinterpret_cast<void*>(static_cast<uintptr_t>(dst_ptrs[d]));
  check_rc(vb200_resize_gather(in_c.data_ptr(), outs, (int)dst_ptrs.size(), dtype_code(in_c.scalar_type(), "resize"), planes, (int)in_h,
                               (int)in_w, (int)out_h, (int)out_w, (int)mode, antialias ? 1 : 0, cur_stream()),
           "resize_gather");
}

// ---- fused inference preprocessing (transforms/_presets.py:57-64) -------------------------------------------------
at::Tensor resize_crop_normalize(const at::Tensor& input, int64_t resize_h, int64_t resize_w, int64_t crop_top, int64_t crop_left,
                                 int64_t crop_h, int64_t crop_w, int64_t mode, bool antialias, at::ArrayRef<double> mean,
                                 at::ArrayRef<double> std) {
  TORCH_CHECK(input.is_cuda(), "input must be a CUDA tensor");
  TORCH_CHECK(input.dim() == 4, "resize_crop_normalize: input must be [B, C, H, W]");
  const int64_t B = input.size(0), C = input.size(1);
  TORCH_CHECK((int64_t)mean.size() == C && (int64_t)std.size() == C && C <= 8, "resize_crop_normalize: one mean / std per channel (<= 8 channels)");
  at::cuda::CUDAGuard guard(input.device());
  at::Tensor in_c = input.contiguous();
  at::Tensor out = at::empty({B, C, crop_h, crop_w}, input.options().dtype(at::kFloat));
  if (out.numel() == 0) return out;
  float m[8], sd[8];
  for (int64_t c = 0; c < C; ++c) { m[c] = (float)mean[c]; sd[c] = (float)std[c]; }
  check_rc(vb200_resize_crop_normalize(in_c.data_ptr(), out.data_ptr<float>(), dtype_code(in_c.scalar_type(), "resize_crop_normalize"), B,
                                       (int)C, (int)input.size(2), (int)input.size(3), (int)resize_h, (int)resize_w, (int)crop_top,
                                       (int)crop_left, (int)crop_h, (int)crop_w, (int)mode, antialias ? 1 : 0, m, sd, cur_stream()),
           "resize_crop_normalize");
  return out;
}

// ---- box_iou_rotated (csrc/ops/box_iou_rotated.cpp; checks as cuda/box_iou_rotated_kernel.cu:92-118) ----------------
at::Tensor box_iou_rotated(const at::Tensor& boxes1, const at::Tensor& boxes2) {
  TORCH_CHECK(boxes1.is_cuda() && boxes2.is_cuda(), "boxes1 and boxes2 must be CUDA tensors");
  TORCH_CHECK(boxes1.dim() == 2 && boxes1.size(1) == 5 && boxes2.dim() == 2 && boxes2.size(1) == 5, "boxes must have shape as Tensor[N, 5]");
  TORCH_CHECK(boxes1.scalar_type() == at::kFloat && boxes2.scalar_type() == at::kFloat, "box_iou_rotated: float32 boxes");
  at::cuda::CUDAGuard guard(boxes1.device());
  at::Tensor b1 = boxes1.contiguous(), b2 = boxes2.contiguous();
  at::Tensor out = at::empty({b1.size(0), b2.size(0)}, b1.options());
  if (out.numel() == 0) return out;
  check_rc(vb200_box_iou_rotated(b1.data_ptr(), b2.data_ptr(), out.data_ptr<float>(), VB200_F32, b1.size(0), b2.size(0), cur_stream()),
           "box_iou_rotated");
  return out;
}

// ---- install / uninstall -----------------------------------------------------
std::unique_ptr<torch::Library> g_override;

void install(bool on) {
  if (!on) { g_override.reset(); return; }
  if (g_override) return;
  auto lib = std::make_unique<torch::Library>(torch::Library::IMPL, "torchvision",
                                              c10::make_optional(c10::DispatchKey::CUDA), __FILE__, __LINE__);
  lib->impl("nms", TORCH_FN(nms));
  lib->impl("roi_align", TORCH_FN(roi_align));
  lib->impl("roi_pool", TORCH_FN(roi_pool));
  lib->impl("ps_roi_align", TORCH_FN(ps_roi_align));
  lib->impl("deform_conv2d", TORCH_FN(deform_conv2d));
  lib->impl("ps_roi_pool", TORCH_FN(ps_roi_pool));
  lib->impl("_ps_roi_pool_backward", TORCH_FN(ps_roi_pool_backward));
  lib->impl("_deform_conv2d_backward", TORCH_FN(deform_conv2d_backward));
  lib->impl("_roi_align_backward", TORCH_FN(roi_align_backward));
  lib->impl("_roi_pool_backward", TORCH_FN(roi_pool_backward));
  lib->impl("_ps_roi_align_backward", TORCH_FN(ps_roi_align_backward));
  g_override = std::move(lib);
}
bool installed() { return (bool)g_override; }
void set_nms_semantics(int64_t s) {
  TORCH_CHECK(s == VB200_NMS_CPU || s == VB200_NMS_CUDA, "nms semantics must be 0 (cpu) or 1 (cuda)");
  g_nms_semantics.store((int)s);
}
int64_t get_nms_semantics() { return g_nms_semantics.load(); }
int64_t launch_count() { return (int64_t)vb200_launch_count(); }
int64_t abi_version() { return vb200_abi_version(); }

}  // namespace

TORCH_LIBRARY(vision_b200, m) {
  m.def("nms(Tensor dets, Tensor scores, float iou_threshold) -> Tensor");
  m.def("batched_nms(Tensor boxes, Tensor scores, Tensor idxs, float iou_threshold) -> Tensor");
  m.def("batched_nms_padded(Tensor boxes, Tensor scores, Tensor idxs, float iou_threshold) -> (Tensor, Tensor)");
  m.def("roi_align(Tensor input, Tensor rois, float spatial_scale, SymInt pooled_height, SymInt pooled_width, int sampling_ratio, bool aligned) -> Tensor");
  m.def("roi_pool(Tensor input, Tensor rois, float spatial_scale, SymInt pooled_height, SymInt pooled_width) -> (Tensor, Tensor)");
  m.def("ps_roi_align(Tensor input, Tensor rois, float spatial_scale, SymInt pooled_height, SymInt pooled_width, int sampling_ratio) -> (Tensor, Tensor)");
  m.def("deform_conv2d(Tensor input, Tensor weight, Tensor offset, Tensor mask, Tensor bias, SymInt stride_h, SymInt stride_w, SymInt pad_h, SymInt pad_w, SymInt dilation_h, SymInt dilation_w, SymInt groups, SymInt offset_groups, bool use_mask) -> Tensor");
  m.def("ps_roi_pool(Tensor input, Tensor rois, float spatial_scale, SymInt pooled_height, SymInt pooled_width) -> (Tensor, Tensor)");
  m.def("_ps_roi_pool_backward(Tensor grad, Tensor rois, Tensor channel_mapping, float spatial_scale, SymInt pooled_height, SymInt pooled_width, SymInt batch_size, SymInt channels, SymInt height, SymInt width) -> Tensor");
  m.def("_deform_conv2d_backward(Tensor grad, Tensor input, Tensor weight, Tensor offset, Tensor mask, Tensor bias, SymInt stride_h, SymInt stride_w, SymInt pad_h, SymInt pad_w, SymInt dilation_h, SymInt dilation_w, SymInt groups, SymInt offset_groups, bool use_mask) -> (Tensor, Tensor, Tensor, Tensor, Tensor)");
  m.def("resize(Tensor input, int out_h, int out_w, int mode, bool antialias) -> Tensor");
  m.def("resize_gather(Tensor input, int[] dst_ptrs, int out_h, int out_w, int mode, bool antialias) -> ()");
  m.def("deform_conv2d_gather(Tensor input, Tensor weight, Tensor offset, Tensor mask, Tensor bias, int[] dst_ptrs, int stride_h, int stride_w, int pad_h, int pad_w, int dilation_h, int dilation_w, int groups, int offset_groups, bool use_mask) -> ()");
  m.def("roi_align_gather(Tensor input, Tensor rois, int[] dst_ptrs, int mc_ptr, float spatial_scale, int pooled_height, int pooled_width, int sampling_ratio, bool aligned) -> ()");
  m.def("resize_crop_normalize(Tensor input, int resize_h, int resize_w, int crop_top, int crop_left, int crop_h, int crop_w, int mode, bool antialias, float[] mean, float[] std) -> Tensor");
  m.def("box_iou_rotated(Tensor boxes1, Tensor boxes2) -> Tensor");
  m.def("detection_postprocess(Tensor boxes, Tensor scores, Tensor labels, float img_h, float img_w, float score_thresh, bool score_inclusive, float min_size, float nms_thresh, int topk) -> (Tensor, Tensor, Tensor)");
  m.def("multiscale_roi_align(Tensor[] features, Tensor rois, float[] scales, int pooled_height, int pooled_width, int sampling_ratio, int k_min, int k_max, float canonical_scale, float canonical_level, float eps) -> (Tensor, Tensor)");
  m.def("_roi_align_backward(Tensor grad, Tensor rois, float spatial_scale, SymInt pooled_height, SymInt pooled_width, SymInt batch_size, SymInt channels, SymInt height, SymInt width, int sampling_ratio, bool aligned) -> Tensor");
  m.def("_roi_pool_backward(Tensor grad, Tensor rois, Tensor argmax, float spatial_scale, SymInt pooled_height, SymInt pooled_width, SymInt batch_size, SymInt channels, SymInt height, SymInt width) -> Tensor");
  m.def("_ps_roi_align_backward(Tensor grad, Tensor rois, Tensor channel_mapping, float spatial_scale, SymInt pooled_height, SymInt pooled_width, int sampling_ratio, SymInt batch_size, SymInt channels, SymInt height, SymInt width) -> Tensor");
  m.def("_install(bool on) -> ()", &install);
  m.def("_installed() -> bool", &installed);
  m.def("_set_nms_semantics(int s) -> ()", &set_nms_semantics);
  m.def("_get_nms_semantics() -> int", &get_nms_semantics);
  m.def("_launch_count() -> int", &launch_count);
  m.def("_abi_version() -> int", &abi_version);
}

TORCH_LIBRARY_IMPL(vision_b200, CUDA, m) {
  m.impl("nms", TORCH_FN(nms));
  m.impl("batched_nms", TORCH_FN(batched_nms));
  m.impl("batched_nms_padded", TORCH_FN(batched_nms_padded));
  m.impl("roi_align", TORCH_FN(roi_align));
  m.impl("roi_pool", TORCH_FN(roi_pool));
  m.impl("ps_roi_align", TORCH_FN(ps_roi_align));
  m.impl("deform_conv2d", TORCH_FN(deform_conv2d));
  m.impl("resize", TORCH_FN(resize));
  m.impl("resize_gather", TORCH_FN(resize_gather));
  m.impl("deform_conv2d_gather", TORCH_FN(deform_conv2d_gather));
  m.impl("roi_align_gather", TORCH_FN(roi_align_gather));
  m.impl("_roi_align_backward", TORCH_FN(roi_align_backward));
  m.impl("_roi_pool_backward", TORCH_FN(roi_pool_backward));
  m.impl("_ps_roi_align_backward", TORCH_FN(ps_roi_align_backward));
  m.impl("multiscale_roi_align", TORCH_FN(multiscale_roi_align));
  m.impl("ps_roi_pool", TORCH_FN(ps_roi_pool));
  m.impl("_ps_roi_pool_backward", TORCH_FN(ps_roi_pool_backward));
  m.impl("detection_postprocess", TORCH_FN(detection_postprocess));
  m.impl("resize_crop_normalize", TORCH_FN(resize_crop_normalize));
  m.impl("box_iou_rotated", TORCH_FN(box_iou_rotated));
  m.impl("_deform_conv2d_backward", TORCH_FN(deform_conv2d_backward));
}

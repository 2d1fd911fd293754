// deform_conv2d_tc.cu — tcgen05 tensor-core path for deform_conv2d (placeholder: not yet enabled).
#include "common.cuh"

namespace vb200 {
struct DcnParams;
int deform_conv2d_tc_try(const void*, const void*, const void*, const void*, const void*, void*, int, const DcnParams&,
                         void*, size_t, cudaStream_t) { return 0; }
size_t deform_conv2d_tc_workspace(int, const DcnParams&) { return 0; }
}  // namespace vb200

// resize_stream.cu — bilinear-antialias downscale fast path (placeholder: not yet enabled).
#include "common.cuh"

namespace vb200 {
int resize_aa_stream_try(const void*, void*, int, int64_t, int, int, int, int, int, cudaStream_t) { return 0; }
}  // namespace vb200

// dcn_params.h — shared between the SIMT and tcgen05 deform_conv2d translation units.
#pragma once
namespace vb200 {
struct DcnParams {
  int batch, c_in, in_h, in_w, c_out, kh, kw, stride_h, stride_w, pad_h, pad_w, dil_h, dil_w;
  int groups, offset_groups, use_mask, out_h, out_w;
};
}  // namespace vb200

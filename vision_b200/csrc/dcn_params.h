// dcn_params.h — shared between the SIMT and tcgen05 deform_conv2d translation units.
#pragma once
namespace vb200 {
struct DcnParams {
  int batch, c_in, in_h, in_w, c_out, kh, kw, stride_h, stride_w, pad_h, pad_w, dil_h, dil_w;
  int groups, offset_groups, use_mask, out_h, out_w;
  int blend16;   // tcgen05 16-bit path: blend the four corners in the storage format (HFMA2) instead of fp32
};
}  // namespace vb200

// TEST INFRASTRUCTURE ONLY — builds oracle/_ref/libbox_iou_rotated_ref.so from the reference's OWN source where it lies:
// /root/reference/torchvision/csrc/ops/box_iou_rotated_utils.h is a self-contained header (cmath / algorithm only), the
// arithmetic core of torchvision::box_iou_rotated (csrc/ops/cpu/box_iou_rotated_kernel.cpp:10-40 loops over it).  Nothing
// is copied into the repo: this file only includes the header by its path under /root/reference and exports a C entry point.
// Used to pin oracle/vision_oracle.c:orc_box_iou_rotated_f32 and to generate tests/golden/box_iou_rotated.npz
// (tests/golden/gen_golden_rotated.py).  The installed torchvision wheel (0.26) does not contain this op.
#include "/root/reference/torchvision/csrc/ops/box_iou_rotated_utils.h"

extern "C" __attribute__((visibility("default"))) void ref_box_iou_rotated_f32(const float* boxes1, int n1, const float* boxes2, int n2,
                                                                                float* ious) {
  // csrc/ops/cpu/box_iou_rotated_kernel.cpp: ious[i * n2 + j] = single_box_iou_rotated<T>(boxes1[i], boxes2[j])
  for (int i = 0; i < n1; ++i)
    for (int j = 0; j < n2; ++j) ious[i * n2 + j] = vision::ops::single_box_iou_rotated<float>(boxes1 + 5 * i, boxes2 + 5 * j);
}

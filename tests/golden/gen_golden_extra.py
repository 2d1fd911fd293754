"""Second fixture file (same rules as gen_golden.py: everything comes from the REFERENCE implementation
imported in the build container): shapes that reach the kernels added late in round 1 —
the 7x7 / sampling_ratio-2 roi_align line kernel, float64 nms, uint8 / fp32 streaming resize.

    python tests/golden/gen_golden_extra.py      # rewrites reference_cpu_extra.npz next to this script
"""
import os

import numpy as np
import torch
import torch.nn.functional as F
import torchvision
from torchvision import ops

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    out = {}
    g = torch.Generator().manual_seed(4321)
    # ---- roi_align, detection-head shape (7x7 bins, sampling_ratio 2, scale 0.25), two images ----
    x = torch.randn(2, 4, 24, 30, generator=g)
    k = 60
    r = torch.zeros(k, 5)
    r[:, 0] = torch.randint(0, 2, (k,), generator=g).float()
    r[:, 1] = torch.rand(k, generator=g) * 120 - 10
    r[:, 2] = torch.rand(k, generator=g) * 96 - 10
    r[:, 3] = r[:, 1] + torch.rand(k, generator=g) * 110 + 1
    r[:, 4] = r[:, 2] + torch.rand(k, generator=g) * 90 + 1
    r[0, 1:] = torch.tensor([-60.0, -60.0, -20.0, -30.0])      # fully outside: zero rows / columns only
    r[1, 1:] = torch.tensor([50.0, 40.0, 50.0, 40.0])          # degenerate
    r[2, 1:] = torch.tensor([100.0, 80.0, 140.0, 110.0])       # hugging the bottom-right border
    out["line_x"], out["line_rois"] = x.numpy(), r.numpy()
    for al in (0, 1):
        out[f"line_out_a{al}"] = ops.roi_align(x, r, (7, 7), 0.25, 2, bool(al)).numpy()
    # ---- float64 nms / batched_nms (cpu/nms_kernel.cpp dispatches float and double) ----
    b = torch.rand(700, 4, generator=g, dtype=torch.float64) * 100
    b[:, 2:] = b[:, :2] + torch.rand(700, 2, generator=g, dtype=torch.float64) * 40 + 0.5
    s = torch.rand(700, generator=g, dtype=torch.float64)
    idx = torch.randint(0, 6, (700,), generator=g)
    out["nms64_boxes"], out["nms64_scores"], out["nms64_idxs"] = b.numpy(), s.numpy(), idx.numpy()
    for i, thr in enumerate((0.3, 0.5)):
        out[f"nms64_keep{i}"] = ops.nms(b, s, thr).numpy()
    out["nms64_thr"] = np.array([0.3, 0.5])
    out["bnms64_keep_v"] = ops.boxes._batched_nms_vanilla(b, s, idx, 0.5).numpy()
    out["bnms64_keep_t"] = ops.boxes._batched_nms_coordinate_trick(b, s, idx, 0.5).numpy()
    # ---- resize of integer / fp32 images the way _geometry.py:340-360 routes CUDA tensors:
    #      to(float32) -> interpolate(antialias) -> round_() for integer dtypes -> to(dtype) ----
    img = torch.randint(0, 256, (1, 3, 48, 512), generator=g, dtype=torch.uint8)
    out["rs8_img"] = img.numpy()
    for size in ((9, 20), (31, 200)):  # scale_w 25.6 and 2.56
        f = F.interpolate(img.float(), size=list(size), mode="bilinear", align_corners=False, antialias=True)
        out[f"rs8_float_{size[0]}x{size[1]}"] = f.numpy()                      # before rounding (ties are visible here)
        out[f"rs8_out_{size[0]}x{size[1]}"] = f.round().to(torch.uint8).numpy()
    imgf = torch.randn(1, 2, 40, 256, generator=g)
    out["rsf_img"] = imgf.numpy()
    out["rsf_out_20x60"] = F.interpolate(imgf, size=[20, 60], mode="bilinear", align_corners=False, antialias=True).numpy()
    out["versions"] = np.array([torch.__version__, torchvision.__version__])
    np.savez_compressed(os.path.join(HERE, "reference_cpu_extra.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_cpu_extra.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()

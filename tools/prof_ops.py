"""Small driver for ncu captures: runs one op of the hot path a few times at its BASELINE shape.
    python tools/prof_ops.py roi_align|roi_pool|ps_roi_align|ps_roi_pool|batched_nms|nms|resize|resize128|resize_noaa|deform|deform_f32|
                             roi_align_bwd|roi_align_bwd_det|multiscale|postprocess|preprocess [iters]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vision_b200 as vb  # noqa: E402
from vision_b200 import workloads  # noqa: E402

vb._lib.load_ops()

op = sys.argv[1]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = "cuda"
if op == "roi_align":
    x, r, kw = workloads.cfg2_roi_align()
    x, r = x.to(dev), r.to(dev)
    fn = lambda: vb.ops.roi_align(x, r, **kw)
elif op == "roi_pool":
    x, r, kw = workloads.cfg2_roi_align()
    x, r = x.to(dev), r.to(dev)
    fn = lambda: vb.ops.roi_pool(x, r, 7, 0.25)
elif op in ("ps_roi_align", "ps_roi_pool"):
    x, r, kw = workloads.cfg2_roi_align()
    x, r = x[:, :245].contiguous().to(dev), r.to(dev)
    fn = (lambda: vb.ops.ps_roi_align(x, r, 7, 0.25, 2)) if op == "ps_roi_align" else (lambda: vb.ops.ps_roi_pool(x, r, 7, 0.25))
elif op in ("roi_align_bwd", "roi_align_bwd_det"):
    _, r, kw = workloads.cfg2_roi_align()
    r = r.to(dev)
    g = torch.randn(1000, 256, 7, 7, device=dev)
    torch.use_deterministic_algorithms(op.endswith("det"))
    fn = lambda: torch.ops.vision_b200._roi_align_backward(g, r, 0.25, 7, 7, 1, 256, 200, 272, 2, False)
elif op == "multiscale":
    from collections import OrderedDict
    gen = torch.Generator().manual_seed(0)
    ih, iw = 800, 1088
    feats = [torch.randn(1, 256, ih // s_, iw // s_, generator=gen).to(dev) for s_ in (4, 8, 16, 32)]
    size = torch.exp(torch.rand(1000, 2, generator=gen) * 4.0 + 2.5)
    xy = torch.rand(1000, 2, generator=gen) * torch.tensor([iw, ih]) * 0.8
    rois = torch.cat([torch.zeros(1000, 1), xy, torch.minimum(xy + size, torch.tensor([float(iw), float(ih)]))], dim=1).to(dev)
    fn = lambda: torch.ops.vision_b200.multiscale_roi_align(feats, rois, [0.25, 0.125, 0.0625, 0.03125], 7, 7, 2, 2, 5, 224.0, 4.0, 1e-6)
elif op == "postprocess":
    from vision_b200 import detection
    gen = torch.Generator().manual_seed(0)
    n = 90_000
    xy = torch.rand(n, 2, generator=gen) * torch.tensor([1000.0, 760.0]); wh = torch.rand(n, 2, generator=gen) * 300 + 1
    bx = torch.cat([xy, xy + wh], 1).to(dev); sc = (torch.rand(n, generator=gen) ** 8).to(dev); lb = (torch.arange(n) % 90).to(dev)
    fn = lambda: detection.detection_postprocess(bx, sc, lb, (800, 1088), 0.05, False, 1e-2, 0.5, 100)
elif op == "preprocess":
    img = torch.randint(0, 256, (64, 3, 500, 375), dtype=torch.uint8, device=dev)
    fn = lambda: vb.transforms.classification_preprocess(img, 224, [256], (0.485, 0.456, 0.406), (0.229, 0.224, 0.225))
elif op == "batched_nms":
    b, s, i = [t.to(dev) for t in workloads.cfg3_batched_nms(clustered=len(sys.argv) > 3)]
    fn = lambda: vb.ops.batched_nms(b, s, i, 0.5)
elif op == "nms":
    b, s, i = [t.to(dev) for t in workloads.cfg3_batched_nms(n=int(os.environ.get('NMS_N', '20000')))]
    fn = lambda: vb.ops.nms(b, s, 0.5)
elif op in ("resize", "resize128", "resize_noaa", "resize_u8", "resize_f32"):
    x = workloads.cfg5_resize(device=dev, batch=128 if op == "resize128" else 32)
    if op == "resize_u8":
        x = (x.float() * 255).round().to(torch.uint8)
    if op == "resize_f32":
        x = x[:16].float()
    fn = lambda: vb.transforms.resize(x, [224, 224], antialias=(op != "resize_noaa"))
elif op in ("deform", "deform_f32"):
    dt = torch.bfloat16 if op == "deform" else torch.float32
    xi, off, w, bi, m = [t.to(dev) for t in workloads.cfg4_deform_conv2d(batch=int(os.environ.get('DCN_BATCH', '32')), dtype=dt)]
    fn = lambda: vb.ops.deform_conv2d(xi, off, w, bi, 1, 1, 1, m)
else:
    raise SystemExit(f"unknown op {op}")
for _ in range(2):
    fn()
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(iters):
    fn()
ev1.record()
torch.cuda.synchronize()
print(f"{op}: {ev0.elapsed_time(ev1) / iters:.4f} ms/iter over {iters} iters (warm L2)")

"""Small driver for ncu captures: runs one op of the hot path a few times at its BASELINE shape.
    python tools/prof_ops.py roi_align|batched_nms|nms|resize|resize_noaa|deform|deform_f32|roi_pool [iters]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vision_b200 as vb  # noqa: E402
from vision_b200 import workloads  # noqa: E402

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
elif op == "batched_nms":
    b, s, i = [t.to(dev) for t in workloads.cfg3_batched_nms(clustered=len(sys.argv) > 3)]
    fn = lambda: vb.ops.batched_nms(b, s, i, 0.5)
elif op == "nms":
    b, s, i = [t.to(dev) for t in workloads.cfg3_batched_nms(n=int(os.environ.get('NMS_N', '20000')))]
    fn = lambda: vb.ops.nms(b, s, 0.5)
elif op in ("resize", "resize_noaa", "resize_u8", "resize_f32"):
    x = workloads.cfg5_resize(device=dev, batch=32)
    if op == "resize_u8":
        x = (x.float() * 255).round().to(torch.uint8)
    if op == "resize_f32":
        x = x[:16].float()
    fn = lambda: vb.transforms.resize(x, [224, 224], antialias=(op != "resize_noaa"))
elif op in ("deform", "deform_f32"):
    dt = torch.bfloat16 if op == "deform" else torch.float32
    xi, off, w, bi, m = [t.to(dev) for t in workloads.cfg4_deform_conv2d(batch=int(os.environ.get('DCN_BATCH', '8')), dtype=dt)]
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

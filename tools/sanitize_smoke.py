"""One small call of every kernel path, meant to run under compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
No numerics are checked here (tests/ does that); the point is out-of-bounds / hazard reports."""
import os
import sys

import torch

sys.path.insert(0, ".")
import vision_b200 as vb
from vision_b200 import workloads

dev = "cuda"
only = sys.argv[1] if len(sys.argv) > 1 else "all"


def env(k, v):
    if v is None:
        os.environ.pop(k, None)
    else:
        os.environ[k] = v
    vb._lib.core().vb200_reload_env()


def roi():
    x, r, kw = workloads.cfg2_roi_align(seed=1, k=300, batch=2, channels=16, height=40, width=52)
    r = r.clone()
    r[::7, 1:3] -= 90.0
    r[1::11, 3:] += 400.0
    x, r = x.to(dev), r.to(dev)
    for path in ("line", "plane", "generic"):
        env("VB200_ROI_ALIGN_PATH", path)
        vb.ops.roi_align(x, r, 7, 0.25, 2, False)
        vb.ops.roi_align(x, r, 7, 0.25, 2, True)
    env("VB200_ROI_ALIGN_PATH", None)
    vb.ops.roi_align(x.half(), r.half(), (3, 5), 0.25, -1, False)
    vb.ops.roi_pool(x, r, 7, 0.25)
    xp = torch.randn(2, 2 * 9, 20, 24, device=dev)
    vb.ops.ps_roi_align(xp, r[:50], 3, 0.25, 2)


def nms():
    for n in (1, 300, 5000):
        b, s, i = [t.to(dev) for t in workloads.cfg3_batched_nms(n=n)]
        for path in ("mask", "chain"):
            env("VB200_NMS_PATH", path)
            vb.ops.nms(b, s, 0.5)
        env("VB200_NMS_PATH", None)
    b, s, i = [t.to(dev) for t in workloads.cfg3_batched_nms(n=30000, classes=40)]
    i[:6000] = 0
    for path in (None, "chain"):
        env("VB200_BNMS_PATH", path)
        vb.ops.batched_nms(b, s, i, 0.5)
    env("VB200_BNMS_PATH", None)
    vb.ops.batched_nms(b[:3000], s[:3000], i[:3000], 0.5)            # coordinate trick
    vb.ops.nms(b[:700].double(), s[:700].double(), 0.5)


def resize():
    for dt in (torch.float16, torch.bfloat16, torch.float32, torch.uint8):
        x = torch.rand(2, 3, 96, 1024, device=dev)
        x = (x * 255).to(torch.uint8) if dt == torch.uint8 else x.to(dt)
        for path in (None, "generic"):
            env("VB200_RESIZE_PATH", path)
            vb.transforms.resize_image(x, [17, 40], antialias=True)
        env("VB200_RESIZE_PATH", None)
        vb.transforms.resize_image(x, [17, 40], antialias=False)
        vb.transforms.resize_image(x, [120, 1100], interpolation="bicubic", antialias=True)


def dcn():
    for dt, cin, cout in ((torch.bfloat16, 64, 128), (torch.float16, 128, 512), (torch.float32, 6, 4)):
        x = torch.randn(2, cin, 12, 12, device=dev).to(dt)
        w = torch.randn(cout, cin, 3, 3, device=dev).to(dt) * 0.05
        off = torch.randn(2, 18, 12, 12, device=dev).to(dt) * 2
        m = torch.rand(2, 9, 12, 12, device=dev).to(dt)
        bias = torch.randn(cout, device=dev).to(dt)
        vb.ops.deform_conv2d(x, off, w, bias, (1, 1), (1, 1), (1, 1), m)
        vb.ops.deform_conv2d(x, off, w, None, (1, 1), (1, 1), (1, 1), None)
        env("VB200_DCN_PATH", "simt")
        vb.ops.deform_conv2d(x, off, w, bias, (1, 1), (1, 1), (1, 1), m)
        env("VB200_DCN_PATH", None)


def band():
    """round 2: the band-resident roi_align kernel (several bands, split bin rows -> RED) and the multi-destination stores"""
    x, r, kw = workloads.cfg2_roi_align(seed=3, k=200, batch=2, channels=8, height=80, width=200)
    r = r.clone()
    r[::7, 1:3] -= 90.0
    r[1::11, 3:] += 400.0
    x, r = x.to(dev), r.to(dev)
    env("VB200_ROI_ALIGN_PATH", "band")
    vb.ops.roi_align(x, r, 7, 0.25, 2, False)
    vb.ops.roi_align(x, r, 7, 0.25, 2, True)
    env("VB200_ROI_ALIGN_PATH", None)
    want = vb.ops.roi_align(x, r, 7, 0.25, 2, False)
    bufs = [torch.empty_like(want) for _ in range(3)]
    torch.ops.vision_b200.roi_align_gather(x, r, [b.data_ptr() for b in bufs], 0, 0.25, 7, 7, 2, False)
    img = torch.rand(2, 3, 96, 1024, device=dev).half()
    o = [torch.empty(2, 3, 17, 40, device=dev, dtype=torch.float16) for _ in range(3)]
    torch.ops.vision_b200.resize_gather(img, [t.data_ptr() for t in o], 17, 40, 0, True)
    xi = torch.randn(2, 64, 12, 12, device=dev).bfloat16()
    w = (torch.randn(128, 64, 3, 3, device=dev) * 0.05).bfloat16()
    off = (torch.randn(2, 18, 12, 12, device=dev) * 2).bfloat16()
    m = torch.rand(2, 9, 12, 12, device=dev).bfloat16()
    bias = torch.randn(128, device=dev).bfloat16()
    d = [torch.empty(2, 128, 12, 12, device=dev, dtype=torch.bfloat16) for _ in range(3)]
    torch.ops.vision_b200.deform_conv2d_gather(xi, w, off, m, bias, [t.data_ptr() for t in d], 1, 1, 1, 1, 1, 1, 1, 1, True)


for name, fn in (("roi", roi), ("nms", nms), ("resize", resize), ("dcn", dcn), ("band", band)):
    if only in ("all", name):
        fn()
        torch.cuda.synchronize()
        print(name, "done", flush=True)

"""Seeded synthetic inputs of the BASELINE.json configurations — shared by
bench.py, the parity tests and __graft_entry__.smoke().  Pure torch; no reference code involved."""
from __future__ import annotations

import torch


def cfg2_roi_align(device="cpu", seed=0, k=1000, batch=1, channels=256, height=200, width=272, dtype=torch.float32):
    """256-ch 200x272 FPN map, 1000 RoIs over the 800x1088 image, scale 0.25, 7x7, sampling_ratio 2."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, channels, height, width, generator=g, dtype=torch.float32)
    img_h, img_w = height * 4, width * 4
    x1 = torch.rand(k, generator=g) * img_w
    y1 = torch.rand(k, generator=g) * img_h
    w = torch.rand(k, generator=g) * (512 - 16) + 16
    h = torch.rand(k, generator=g) * (512 - 16) + 16
    rois = torch.stack([torch.randint(0, batch, (k,), generator=g).float(), x1, y1,
                        (x1 + w).clamp(max=img_w), (y1 + h).clamp(max=img_h)], dim=1)
    return x.to(device=device, dtype=dtype), rois.to(device=device, dtype=dtype), dict(
        output_size=(7, 7), spatial_scale=0.25, sampling_ratio=2, aligned=False)


def cfg3_batched_nms(device="cpu", seed=0, n=100_000, classes=80, clustered=False):
    """100k boxes x 80 classes, distinct scores (the reference's final sort is unstable, boxes.py:126)."""
    g = torch.Generator().manual_seed(seed)
    if clustered:
        centres = n // 50
        cxy = torch.rand(centres, 2, generator=g) * 1000
        cwh = torch.rand(centres, 2, generator=g) * 200 + 1
        rep = torch.arange(n) % centres
        xy = cxy[rep] + torch.randn(n, 2, generator=g) * 0.05 * cwh[rep]
        wh = cwh[rep] * (1 + torch.randn(n, 2, generator=g) * 0.05).clamp(min=0.5)
    else:
        xy = torch.rand(n, 2, generator=g) * 1000
        wh = torch.rand(n, 2, generator=g) * 200 + 1
    boxes = torch.cat([xy, xy + wh], dim=1).float()
    scores = torch.randperm(n, generator=g).float() / n
    idxs = torch.randint(0, classes, (n,), generator=g)
    return boxes.to(device), scores.to(device), idxs.to(device)


def cfg4_deform_conv2d(device="cpu", seed=0, batch=32, c_in=512, c_out=512, hw=64, dtype=torch.bfloat16,
                       offset_scale=2.0, use_mask=True):
    """3x3 DCNv2, stride 1 pad 1: values are rounded to `dtype` (the reference is run in fp32 on them)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, c_in, hw, hw, generator=g)
    w = torch.randn(c_out, c_in, 3, 3, generator=g) * (1.0 / (c_in * 9) ** 0.5)
    off = torch.randn(batch, 18, hw, hw, generator=g) * offset_scale
    mask = torch.rand(batch, 9, hw, hw, generator=g) if use_mask else None
    bias = torch.randn(c_out, generator=g)
    cast = lambda t: None if t is None else t.to(dtype).to(device)
    return cast(x), cast(off), cast(w), cast(bias), cast(mask)


def cfg5_resize(device="cpu", seed=0, batch=1024, dtype=torch.float16, height=2160, width=3840):
    """[batch,3,2160,3840] -> 224x224; generated on `device` (51 GB at batch 1024)."""
    g = torch.Generator(device=device).manual_seed(seed)
    return torch.rand(batch, 3, height, width, generator=g, device=device, dtype=torch.float32).to(dtype) \
        if batch <= 8 else torch.rand(batch, 3, height, width, generator=g, device=device, dtype=dtype)

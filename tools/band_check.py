"""roi_align path A/B on the GPU: band-resident vs line kernel on cfg2 (CUDA events, L2 flushed), plus agreement."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vision_b200 as vb  # noqa: E402
from vision_b200 import _lib, workloads  # noqa: E402

DEV = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def timed(fn, iters=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def setenv(**kv):
    for k, v in kv.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)
    _lib.core().vb200_reload_env()


def main():
    x, rois, kw = workloads.cfg2_roi_align()
    xd, rd = x.to(DEV), rois.to(DEV)
    outs = {}
    for path in ("line", "band"):
        setenv(VB200_ROI_ALIGN_PATH=path)
        outs[path] = vb.ops.roi_align(xd, rd, **kw)
        med, best = timed(lambda: vb.ops.roi_align(xd, rd, **kw))
        print(f"cfg2 {path}: median {med:.1f} us  best {best:.1f} us", flush=True)
    err = (outs["band"] - outs["line"]).abs().max().item()
    print("max |band - line| =", err, flush=True)
    setenv(VB200_ROI_ALIGN_PATH="band")
    for ovh in sys.argv[1:]:
        setenv(VB200_ROI_BAND_OVH=ovh)
        med, best = timed(lambda: vb.ops.roi_align(xd, rd, **kw))
        print(f"cfg2 band ovh={ovh}: median {med:.1f} us  best {best:.1f} us", flush=True)
    setenv(VB200_ROI_BAND_OVH=None)
    # other shapes: batch 2, 512 RoIs per image; 4000 RoIs
    for (b, k) in ((2, 1000), (1, 4000), (4, 2000)):
        x2, r2, _ = workloads.cfg2_roi_align(seed=1, k=k, batch=b)
        x2, r2 = x2.to(DEV), r2.to(DEV)
        res = {}
        for path in ("line", "band"):
            setenv(VB200_ROI_ALIGN_PATH=path)
            res[path] = timed(lambda: vb.ops.roi_align(x2, r2, **kw), iters=10, warm=2)[0]
        print(f"batch {b} rois {k}: line {res['line']:.1f} us, band {res['band']:.1f} us", flush=True)


if __name__ == "__main__":
    main()

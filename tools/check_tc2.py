"""GPU check of the CTA-pair (cta_group::2) deform_conv2d kernel against the single-CTA tcgen05 kernel.
    python tools/check_tc2.py            (run under `timeout`: a protocol bug in a new kernel shows up as a hang)"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vision_b200 as vb  # noqa: E402
from vision_b200 import workloads  # noqa: E402


def run(cases):
    for (batch, cin, cout, hw, dt) in cases:
        xi, off, w, bi, m = [t.cuda() for t in workloads.cfg4_deform_conv2d(seed=batch + hw, batch=batch, c_in=cin, c_out=cout, hw=hw, dtype=dt)]
        os.environ["VB200_DCN_CTA2"] = "0"
        vb._lib.core().vb200_reload_env()
        ref = vb.ops.deform_conv2d(xi, off, w, bi, 1, 1, 1, m)
        torch.cuda.synchronize()
        os.environ["VB200_DCN_CTA2"] = "1"
        vb._lib.core().vb200_reload_env()
        t0 = time.time()
        got = vb.ops.deform_conv2d(xi, off, w, bi, 1, 1, 1, m)
        torch.cuda.synchronize()
        err = (got.float() - ref.float()).abs().max().item()
        print(f"batch={batch} cin={cin} cout={cout} hw={hw} {dt}: max|cta2 - cta1| = {err:.3e}  ({time.time() - t0:.3f}s)", flush=True)
        assert err <= 2e-2, err


if __name__ == "__main__":
    run([(1, 64, 512, 12, torch.bfloat16), (3, 128, 512, 20, torch.bfloat16), (2, 128, 1024, 16, torch.float16),
         (4, 512, 512, 64, torch.bfloat16)])
    print("tc2 ok")

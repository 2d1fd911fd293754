#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (driver contract: one JSON line on stdout).

Workload (BASELINE.json configs[1]): torchvision.ops.roi_align, 256-ch 200x272 fp32 FPN feature
map, 1000 RoIs, 7x7 output, spatial_scale 0.25, sampling_ratio 2.  A "step" = one roi_align call
over one such batch, through the reference-facing API (torchvision.ops.roi_align after
vision_b200.install() -> dispatcher -> C ABI -> sm_100a kernels).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  torchrun --nproc-per-node N ... bench.py --gpus N ...      (one rank per GPU, NCCL)

N > 1 is weak scaling: every rank owns one image (feature map + its 1000 RoIs).  Images are
independent units, so the timed step has NO data-path collective (tier rule 5); the north-star's single
NCCL all-gather of the per-shard outputs is measured separately and reported under "with_allgather".

Timing: W >= 3 warm-up steps; L2 is flushed (256 MiB write) before every timed step; each step is
bracketed by CUDA events on the launching stream and the K step times are summed; barrier +
synchronize on both sides; max over ranks.  `--impl reference` times the reference's own CPU kernel
(installed torchvision wheel; the oracle port if it is absent) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ALG_BYTES = 55_705_600 + 20_000 + 50_176_000      # map + rois + output
K_ROIS = 1000
# dram__bytes_read.sum + dram__bytes_write.sum of roi_align_line_kernel<7, 2>, one launch of this workload,
# from profiles/r1_roi_align_line_v3.ncu-rep (67,274,240 + 9,228,288 B; most of the 50 MB output is still
# dirty in the 126 MB L2 when the capture ends)
NCU_TRAFFIC_BYTES = 67_274_240 + 9_228_288
WORKLOAD = "roi_align fp32 1x256x200x272, 1000 RoIs, 7x7, scale 0.25, sampling_ratio 2, aligned=False (BASELINE configs[1])"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_fn():
    """The reference's own CPU implementation of the path, if the wheel is importable; else the oracle port.
    Returns (fn, kind, description, threads).  The reference kernel is a single-threaded loop over RoIs
    (csrc/ops/cpu/roi_align_kernel.cpp:33-35), so "all the host threads it can use" is one thread per call; to
    give it the whole machine the RoIs are split into one chunk per core and the UNMODIFIED op is called on the
    chunks from a thread pool (the op releases the GIL; outputs are concatenated in RoI order)."""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    from vision_b200 import workloads

    x, rois, kw = workloads.cfg2_roi_align()
    threads = max(1, min(os.cpu_count() or 1, 64, rois.shape[0]))
    chunks = [c for c in torch.chunk(rois, threads, dim=0) if c.shape[0]]
    pool = ThreadPoolExecutor(max_workers=len(chunks))
    try:
        import torchvision

        def fn():
            return torch.cat(list(pool.map(lambda r: torchvision.ops.roi_align(x, r, **kw), chunks)), dim=0)

        return fn, "reference", (f"torchvision {torchvision.__version__} CPU kernel (csrc/ops/cpu/roi_align_kernel.cpp, a "
                                 f"single-threaded loop) called on {len(chunks)} RoI chunks from {len(chunks)} threads"), len(chunks)
    except Exception:
        import numpy as np
        import oracle

        xn = x.numpy()
        rn = [c.numpy() for c in chunks]

        def fn():
            return np.concatenate(list(pool.map(lambda r: oracle.roi_align(xn, r, kw["output_size"], kw["spatial_scale"],
                                                                         kw["sampling_ratio"], kw["aligned"]), rn)), axis=0)

        return fn, "port", f"oracle/vision_oracle.c restatement (single-threaded C) on {len(chunks)} RoI chunks / threads", len(chunks)


def time_cpu(fn, calls: int) -> float:
    fn()
    t0 = time.perf_counter()
    for _ in range(calls):
        fn()
    return (time.perf_counter() - t0) / calls


def run_reference(args, rank: int):
    if rank != 0:
        return
    import torch

    torch.set_num_threads(1)          # the pool supplies the parallelism; no intra-op threads under it
    fn, kind, desc, threads = cpu_reference_fn()
    for _ in range(min(args.warmup, 2)):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = time.perf_counter() - t0
    ms = dt / args.steps * 1e3
    val = K_ROIS / (ms / 1e3)
    line = {
        "impl": "reference", "metric": "roi_align RoIs/s", "value": val, "unit": "RoIs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "l2": "n/a (CPU)", "parallelism": "host cores"},
        "cpu_baseline": {"value": val, "unit": "RoIs/s", "cores": threads, "host_cores": os.cpu_count(), "kind": kind,
                         "sample": f"{args.steps} full-size calls of the workload; {desc}"},
        "e2e": {"value": val, "unit": "RoIs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def secondary_numbers(torch, vb, dev) -> dict:
    """Quick device-timed numbers for the other BASELINE configs (not the headline; reduced batch where stated)."""
    from vision_b200 import workloads
    out = {}

    def timed(fn, iters):
        fn(); torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            fn()
        e.record(); torch.cuda.synchronize()
        return s.elapsed_time(e) / iters

    try:
        b, sc, ix = [t.to(dev) for t in workloads.cfg3_batched_nms()]
        ms = timed(lambda: vb.ops.batched_nms(b, sc, ix, 0.5), 10)
        out["batched_nms_100k_x80"] = {"ms": ms, "boxes_per_s": 100_000 / (ms / 1e3), "note": "includes the output-size sync"}
        del b, sc, ix
        nb = 128                                            # the per-GPU shard of cfg5 at 8 GPUs (6.4 GB)
        x = workloads.cfg5_resize(device=dev, batch=nb)
        ms = timed(lambda: vb.transforms.resize(x, [224, 224]), 5)
        nbytes = x.numel() * 2 + nb * 3 * 224 * 224 * 2
        peak, _ = peaks()
        out["resize_fp16_2160x3840_to_224_batch128"] = {"ms": ms, "images_per_s": nb / (ms / 1e3), "GBps": nbytes / ms / 1e6,
                                                        "hbm_frac": nbytes / ms / 1e6 / peak, "algorithmic_bytes": nbytes}
        del x
        torch.cuda.empty_cache()
        xi, off, w, bi, m = [t.to(dev) for t in workloads.cfg4_deform_conv2d(batch=32)]
        ms = timed(lambda: vb.ops.deform_conv2d(xi, off, w, bi, 1, 1, 1, m), 5)
        fl = 2 * 32 * 64 * 64 * 512 * 512 * 9
        tf_peak = 1461.6
        try:
            tf_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
        except Exception:
            pass
        out["deform_conv2d_bf16_cfg4_full"] = {"ms": ms, "TFLOPs": fl / ms / 1e9, "tensor_frac_of_sustained_peak": fl / ms / 1e9 / tf_peak,
                                               "note": "whole op: NCHW->NHWC staging + weight packing + tcgen05 kernel"}
    except Exception as ex:   # secondary numbers never fail the headline
        out["error"] = repr(ex)[:300]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--cpu-calls", type=int, default=10, help="CPU-baseline sample size (full-size calls)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    assert args.warmup >= 3, "timing rules: W >= 3"
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback on the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import torchvision
    import vision_b200 as vb
    from vision_b200 import sharded, workloads

    vb.install()
    x, rois, kw = workloads.cfg2_roi_align(seed=rank)
    xd, rd = x.to(dev), rois.to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def step():
        return torchvision.ops.roi_align(xd, rd, **kw)

    def step_gather():
        return sharded.all_gather_equal(torchvision.ops.roi_align(xd, rd, **kw))

    for _ in range(args.warmup):
        flush.zero_()
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    launches0 = vb.launch_count()
    wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()                       # L2 flush between timed iterations (not timed)
        starts[i].record(stream)
        step()
        ends[i].record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    launches = vb.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
    tt = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms = float(tt.item())
    ms_per_step = total_ms / args.steps

    # ---- end to end: pinned host buffers; EVERY step copies its inputs H2D and its result D2H ----
    # Three streams (copy-in / compute / copy-out) with double buffers, so step i's D2H overlaps step
    # i+1's H2D (PCIe is full duplex); all copies stay inside the timed region.
    xh, rh = x.pin_memory(), rois.pin_memory()
    oh = [torch.empty(K_ROIS, 256, 7, 7, dtype=torch.float32).pin_memory() for _ in range(2)]
    xdb = [torch.empty_like(xd) for _ in range(2)]
    rdb = [torch.empty_like(rd) for _ in range(2)]
    s_in, s_cmp, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_cmp = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]
    e_steps = max(6, min(args.steps, 20))

    def e2e_run(n):
        outs = [None, None]
        for i in range(n):
            bsel = i % 2
            with torch.cuda.stream(s_in):
                if i >= 2:
                    s_in.wait_event(ev_cmp[bsel])          # device input buffer free again
                xdb[bsel].copy_(xh, non_blocking=True)
                rdb[bsel].copy_(rh, non_blocking=True)
                ev_in[bsel].record(s_in)
            with torch.cuda.stream(s_cmp):
                s_cmp.wait_event(ev_in[bsel])
                o = torchvision.ops.roi_align(xdb[bsel], rdb[bsel], **kw)
                ev_cmp[bsel].record(s_cmp)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_cmp[bsel])
                o.record_stream(s_out)
                oh[bsel].copy_(o, non_blocking=True)
                ev_out[bsel].record(s_out)
            outs[bsel] = o

    torch.cuda.synchronize()
    e2e_run(4)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(s_in)
    e2e_run(e_steps)
    e.record(s_out)
    torch.cuda.synchronize()
    e2e_ms = torch.tensor([s.elapsed_time(e) / e_steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_ms.item())

    # ---- the optional single all-gather of per-shard outputs (N > 1), timed separately ----
    gather_ms = None
    if world > 1:
        for _ in range(3):
            step_gather()
        torch.cuda.synchronize(); dist.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g_steps = max(5, min(args.steps, 20))
        g0.record(stream)
        for _ in range(g_steps):
            step_gather()
        g1.record(stream)
        torch.cuda.synchronize(); dist.barrier()
        gt = torch.tensor([g0.elapsed_time(g1) / g_steps], device=dev, dtype=torch.float64)
        dist.all_reduce(gt, op=dist.ReduceOp.MAX)
        gather_ms = float(gt.item())

    if rank == 0:
        peak, peak_src = peaks()
        achieved = ALG_BYTES / (ms_per_step / 1e3) / 1e9     # per GPU (every rank runs the same kernel on its own image)
        # CPU baseline: bounded sample on this box's host cores (rank 0, N=1 only)
        cpu = None
        if world == 1:
            torch.set_num_threads(1)      # the pool supplies the parallelism
            fn, kind, desc, threads = cpu_reference_fn()
            sec = time_cpu(fn, args.cpu_calls)
            cpu = {"value": K_ROIS / sec, "unit": "RoIs/s", "cores": threads, "host_cores": os.cpu_count(), "kind": kind,
                   "sample": f"{args.cpu_calls} full-size calls ({sec * 1e3:.0f} ms each); {desc}"}
        line = {
            "metric": "roi_align RoIs/s", "value": world * K_ROIS / (ms_per_step / 1e3), "unit": "RoIs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "l2": "flushed before every timed step (256 MiB write); per-step CUDA events summed",
                       "parallelism": f"dp{world}: one image per rank, no data-path collective in the timed step",
                       "api": "torchvision.ops.roi_align after vision_b200.install()"},
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": NCU_TRAFFIC_BYTES,
                "traffic_source": "profiles/r1_roi_align_line_v3.ncu-rep (ncu --set full, one launch)",
                "kernel": "roi_align_line_kernel<7, 2> (+ ~4 us roi_align_line_geometry_kernel inside the same event pair)",
                "algorithmic_bytes": ALG_BYTES, "peak_source": peak_src},
            "cpu_baseline": cpu,
            "e2e": {"value": world * K_ROIS / (e2e_ms / 1e3), "unit": "RoIs/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": x.numel() * 4 + rois.numel() * 4, "d2h_bytes_per_step": oh[0].numel() * 4,
                    "note": "pinned host buffers, H2D + op + D2H every step; 3 streams, double-buffered"},
            "gpu_launches": int(launches), "clocks": clocks, "wall_s_timed_region": wall,
        }
        if gather_ms is not None:
            line["with_allgather"] = {"ms_per_step": gather_ms, "value": world * K_ROIS / (gather_ms / 1e3), "unit": "RoIs/s",
                                      "bytes_gathered_per_rank": world * K_ROIS * 256 * 49 * 4,
                                      "note": "op + ONE all_gather_into_tensor of the per-shard outputs (NCCL), L2 not flushed"}
        if world == 1 and not args.no_secondary:
            line["secondary"] = secondary_numbers(torch, vb, dev)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

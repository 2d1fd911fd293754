"""CPU oracle — TEST INFRASTRUCTURE ONLY (see oracle/vision_oracle.c header).

numpy-in / numpy-out ctypes bindings over ``libvision_oracle.so``.  Only
``tests/``, ``bench.py``'s cpu_baseline leg and ``__graft_entry__.smoke()`` may
import this package; ``vision_b200`` (the product) never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libvision_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "vision_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.orc_nms_f32.restype = ctypes.c_int64
        _lib.orc_batched_nms_f32.restype = ctypes.c_int64
        _lib.orc_nms_f64.restype = ctypes.c_int64
        _lib.orc_batched_nms_f64.restype = ctypes.c_int64
        _lib.orc_deform_conv2d_f32.restype = ctypes.c_int
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


NMS_MODE_CPU = 0   # arithmetic of csrc/ops/cpu/nms_kernel.cpp
NMS_MODE_CUDA = 1  # arithmetic of the compiled csrc/ops/cuda/nms_kernel.cu (FMA-contracted Sa+Sb, float thr)
NMS_MODE_CUDA_HALF = 2  # the compiled devIoU<Half>: pass fp16 values (any float container); pinned on the GPU box only


def _is_f64(a) -> bool:
    return getattr(a, "dtype", None) == np.float64


def nms(boxes, scores, iou_threshold: float, mode: int = NMS_MODE_CPU) -> np.ndarray:
    """float64 inputs use the double twin (the reference dispatches nms on float and double)."""
    f64 = _is_f64(boxes)
    cast = (lambda a: np.ascontiguousarray(a, dtype=np.float64)) if f64 else _f32
    boxes, scores = cast(boxes).reshape(-1, 4), cast(scores).reshape(-1)
    n = boxes.shape[0]
    keep = np.empty(max(n, 1), dtype=np.int64)
    fn = lib().orc_nms_f64 if f64 else lib().orc_nms_f32
    k = fn(_p(boxes), _p(scores), ctypes.c_int64(n), ctypes.c_double(iou_threshold),
                          ctypes.c_int(mode), _p(keep))
    return keep[:k].copy()


def batched_nms(boxes, scores, idxs, iou_threshold: float, mode: int = NMS_MODE_CPU,
                strategy: int = 0, device_is_cuda: bool = False) -> np.ndarray:
    """strategy: 0 = reference's own switch, 1 = vanilla, 2 = coordinate trick."""
    f64 = _is_f64(boxes)
    cast = (lambda a: np.ascontiguousarray(a, dtype=np.float64)) if f64 else _f32
    boxes, scores = cast(boxes).reshape(-1, 4), cast(scores).reshape(-1)
    idxs = np.ascontiguousarray(idxs, dtype=np.int64).reshape(-1)
    n = boxes.shape[0]
    keep = np.empty(max(n, 1), dtype=np.int64)
    fn = lib().orc_batched_nms_f64 if f64 else lib().orc_batched_nms_f32
    k = fn(_p(boxes), _p(scores), _p(idxs), ctypes.c_int64(n),
                                  ctypes.c_double(iou_threshold), ctypes.c_int(mode),
                                  ctypes.c_int(strategy), ctypes.c_int(int(device_is_cuda)), _p(keep))
    return keep[:k].copy()


def roi_align(inp, rois, output_size, spatial_scale=1.0, sampling_ratio=-1, aligned=False) -> np.ndarray:
    inp, rois = _f32(inp), _f32(rois).reshape(-1, 5)
    _, c, h, w = inp.shape
    ph, pw = (output_size, output_size) if isinstance(output_size, int) else output_size
    k = rois.shape[0]
    out = np.zeros((k, c, ph, pw), dtype=np.float32)
    lib().orc_roi_align_f32(_p(inp), _p(rois), c, h, w, k, ph, pw, ctypes.c_float(spatial_scale),
                            int(sampling_ratio), int(bool(aligned)), _p(out))
    return out


def roi_pool(inp, rois, output_size, spatial_scale=1.0):
    inp, rois = _f32(inp), _f32(rois).reshape(-1, 5)
    _, c, h, w = inp.shape
    ph, pw = (output_size, output_size) if isinstance(output_size, int) else output_size
    k = rois.shape[0]
    out = np.zeros((k, c, ph, pw), dtype=np.float32)
    arg = np.zeros((k, c, ph, pw), dtype=np.int32)
    lib().orc_roi_pool_f32(_p(inp), _p(rois), c, h, w, k, ph, pw, ctypes.c_float(spatial_scale),
                           _p(out), _p(arg))
    return out, arg


def ps_roi_align(inp, rois, output_size, spatial_scale=1.0, sampling_ratio=-1):
    inp, rois = _f32(inp), _f32(rois).reshape(-1, 5)
    _, c, h, w = inp.shape
    ph, pw = (output_size, output_size) if isinstance(output_size, int) else output_size
    assert c % (ph * pw) == 0
    k = rois.shape[0]
    co = c // (ph * pw)
    out = np.zeros((k, co, ph, pw), dtype=np.float32)
    cm = np.zeros((k, co, ph, pw), dtype=np.int32)
    lib().orc_ps_roi_align_f32(_p(inp), _p(rois), c, h, w, k, ph, pw, ctypes.c_float(spatial_scale),
                               int(sampling_ratio), _p(out), _p(cm))
    return out, cm


def ps_roi_pool(inp, rois, output_size, spatial_scale=1.0):
    inp, rois = _f32(inp), _f32(rois).reshape(-1, 5)
    _, c, h, w = inp.shape
    ph, pw = (output_size, output_size) if isinstance(output_size, int) else output_size
    assert c % (ph * pw) == 0
    k = rois.shape[0]
    co = c // (ph * pw)
    out = np.zeros((k, co, ph, pw), dtype=np.float32)
    cm = np.zeros((k, co, ph, pw), dtype=np.int32)
    lib().orc_ps_roi_pool_f32(_p(inp), _p(rois), c, h, w, k, ph, pw, ctypes.c_float(spatial_scale), _p(out), _p(cm))
    return out, cm


def box_iou_rotated(boxes1, boxes2) -> np.ndarray:
    """[n1, 5] x [n2, 5] boxes (x_ctr, y_ctr, w, h, angle in degrees) -> [n1, n2] IoU."""
    b1, b2 = _f32(boxes1).reshape(-1, 5), _f32(boxes2).reshape(-1, 5)
    out = np.zeros((b1.shape[0], b2.shape[0]), dtype=np.float32)
    lib().orc_box_iou_rotated_f32(_p(b1), b1.shape[0], _p(b2), b2.shape[0], _p(out))
    return out


def box_iou_rotated_ref(boxes1, boxes2):
    """The reference's own arithmetic (oracle/_ref, compiled from /root/reference's header); None when it was not built."""
    path = os.path.join(_HERE, "_ref", "libbox_iou_rotated_ref.so")
    if not os.path.exists(path):
        return None
    ref = ctypes.CDLL(path)
    b1, b2 = _f32(boxes1).reshape(-1, 5), _f32(boxes2).reshape(-1, 5)
    out = np.zeros((b1.shape[0], b2.shape[0]), dtype=np.float32)
    ref.ref_box_iou_rotated_f32(_p(b1), b1.shape[0], _p(b2), b2.shape[0], _p(out))
    return out


def deform_conv2d(inp, offset, weight, bias=None, stride=(1, 1), padding=(0, 0), dilation=(1, 1),
                  mask=None) -> np.ndarray:
    inp, offset, weight = _f32(inp), _f32(offset), _f32(weight)
    b, cin, ih, iw = inp.shape
    cout, cin_g, kh, kw = weight.shape
    sh, sw = stride
    ph, pw = padding
    dh, dw = dilation
    groups = cin // cin_g
    off_groups = offset.shape[1] // (2 * kh * kw)
    oh = (ih + 2 * ph - (dh * (kh - 1) + 1)) // sh + 1
    ow = (iw + 2 * pw - (dw * (kw - 1) + 1)) // sw + 1
    out = np.zeros((b, cout, oh, ow), dtype=np.float32)
    use_mask = mask is not None
    mask_a = _f32(mask) if use_mask else np.zeros(1, np.float32)
    bias_a = _f32(bias) if bias is not None else np.zeros(cout, np.float32)
    if b == 0:
        return out
    rc = lib().orc_deform_conv2d_f32(_p(inp), _p(weight), _p(offset), _p(mask_a), _p(bias_a), b, cin,
                                     ih, iw, cout, kh, kw, sh, sw, ph, pw, dh, dw, groups, off_groups,
                                     int(use_mask), _p(out))
    assert rc == 0
    return out


RESIZE_BILINEAR, RESIZE_BICUBIC = 0, 1


def resize(inp, out_hw, mode: int = RESIZE_BILINEAR, antialias: bool = True) -> np.ndarray:
    """inp [..., H, W] float32 planes -> [..., OH, OW] (compute in fp32, like the reference's
    fp16->fp32->interpolate->fp16 route, _geometry.py:340-360; casts are the caller's)."""
    inp = _f32(inp)
    lead, (h, w) = inp.shape[:-2], inp.shape[-2:]
    oh, ow = out_hw
    planes = int(np.prod(lead)) if lead else 1
    out = np.zeros(lead + (oh, ow), dtype=np.float32)
    lib().orc_resize_f32(_p(inp), ctypes.c_int64(planes), h, w, oh, ow, int(mode), int(bool(antialias)),
                         _p(out))
    return out

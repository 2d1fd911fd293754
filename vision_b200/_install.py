"""install()/uninstall(): put the kernels behind torchvision's own API surface.

1. dispatcher ops (nms, roi_align, roi_pool, ps_roi_align, deform_conv2d): re-registered for the
   CUDA key of the existing ``torchvision::`` schemas by the C++ shim (a run-time twin of the
   reference's TORCH_LIBRARY_IMPL blocks, e.g. csrc/ops/cuda/roi_align_kernel.cu:470-477).  Meta,
   Autograd, Autocast, CPU and quantized registrations are untouched.
2. ``batched_nms`` is Python in the reference (torchvision/ops/boxes.py:57-126): the module
   attribute is rebound (detection models call ``box_ops.batched_nms`` at call time).
3. ``MultiScaleRoIAlign`` (torchvision/ops/poolers.py:147-228): the module-level ``_multiscale_roi_align`` is rebound
   to the fused kernel (device-side LevelMapper + one gather launch over all FPN levels) when the shape is covered.
4. detection post-processing: ``RoIHeads.postprocess_detections`` and ``RegionProposalNetwork.filter_proposals`` keep their
   tensor prologue and run the per-image tail (clip, filters, batched_nms, top-k, gathers) as one fused call.
5. ``resize`` has no torchvision kernel (transforms/v2/functional/_geometry.py:283-362 calls
   F.interpolate): the entries of ``_KERNEL_REGISTRY[resize]`` for Tensor / Image / Video are swapped.
CPU tensors and unsupported dtypes/modes keep flowing to the reference implementation.
"""
from __future__ import annotations

import functools
import warnings

import torch

from . import _lib, transforms as _tf

_state: dict = {}


def installed() -> bool:
    return bool(_state)


def install() -> None:
    if _state:
        return
    import torchvision  # the schemas must exist before the CUDA key is overridden
    from torchvision.ops import boxes as tv_boxes
    from torchvision.transforms.v2.functional import _geometry as tv_geo, _utils as tv_utils
    from torchvision import tv_tensors

    _lib.load_ops()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")   # "Overriding a previously registered kernel" (expected)
        torch.ops.vision_b200._install(True)

    # ---- batched_nms ----
    orig_batched_nms = tv_boxes.batched_nms

    @functools.wraps(orig_batched_nms)
    def batched_nms(boxes, scores, idxs, iou_threshold):
        if (isinstance(boxes, torch.Tensor) and boxes.is_cuda and boxes.dtype in (torch.float32, torch.float64, torch.float16)
                and not torch.jit.is_scripting() and not torch.jit.is_tracing()):
            return torch.ops.vision_b200.batched_nms(boxes, scores, idxs, float(iou_threshold))
        return orig_batched_nms(boxes, scores, idxs, iou_threshold)

    tv_boxes.batched_nms = batched_nms
    torchvision.ops.batched_nms = batched_nms

    # ---- MultiScaleRoIAlign: the per-level loop of poolers.py:147-228 becomes one fused call when the kernel covers it ----
    from torchvision.ops import poolers as tv_poolers
    from . import ops as _ops_mod

    orig_msra = tv_poolers._multiscale_roi_align

    def _multiscale_roi_align(x_filtered, boxes, output_size, sampling_ratio, scales, mapper):
        if (scales is not None and mapper is not None and not torch.jit.is_scripting() and not torch.jit.is_tracing()
                and not torchvision._is_tracing() and _ops_mod.multiscale_roi_align_supported(x_filtered, boxes, output_size, sampling_ratio)):
            return _ops_mod.multiscale_roi_align(x_filtered, boxes, output_size, sampling_ratio, scales, mapper)
        return orig_msra(x_filtered, boxes, output_size, sampling_ratio, scales, mapper)

    tv_poolers._multiscale_roi_align = _multiscale_roi_align

    # ---- detection post-processing around batched_nms (roi_heads.py:680-737, rpn.py:242-298) ----
    from torchvision.models.detection import roi_heads as tv_roi_heads, rpn as tv_rpn
    from . import detection as _det

    orig_pp = tv_roi_heads.RoIHeads.postprocess_detections
    orig_fp = tv_rpn.RegionProposalNetwork.filter_proposals

    def postprocess_detections(self, class_logits, box_regression, proposals, image_shapes):
        if _det._fusable(class_logits) and not torchvision._is_tracing():
            return _det.roi_heads_postprocess_detections(self, class_logits, box_regression, proposals, image_shapes, _orig=orig_pp)
        return orig_pp(self, class_logits, box_regression, proposals, image_shapes)

    def filter_proposals(self, proposals, objectness, image_shapes, num_anchors_per_level):
        if _det._fusable(proposals) and not torchvision._is_tracing():
            return _det.rpn_filter_proposals(self, proposals, objectness, image_shapes, num_anchors_per_level, _orig=orig_fp)
        return orig_fp(self, proposals, objectness, image_shapes, num_anchors_per_level)

    tv_roi_heads.RoIHeads.postprocess_detections = postprocess_detections
    tv_rpn.RegionProposalNetwork.filter_proposals = filter_proposals

    # ---- ImageClassification preset (transforms/_presets.py:57-64): resize + center_crop + to float + normalize fused ----
    from torchvision.transforms import _presets as tv_presets

    orig_preset_forward = tv_presets.ImageClassification.forward

    def preset_forward(self, img):
        if _tf.classification_preprocess_supported(img, self.crop_size, self.resize_size, self.interpolation, self.antialias):
            return _tf.classification_preprocess(img, self.crop_size, self.resize_size, self.mean, self.std, self.interpolation, self.antialias)
        return orig_preset_forward(self, img)

    tv_presets.ImageClassification.forward = preset_forward

    # ---- resize ----
    registry = tv_utils._KERNEL_REGISTRY[tv_geo.resize]
    saved = dict(registry)
    orig_image = tv_geo.resize_image

    @functools.wraps(orig_image)
    def resize_image(image, size, interpolation=tv_geo.InterpolationMode.BILINEAR, max_size=None, antialias=True):
        if isinstance(image, torch.Tensor) and _tf.supports(image, interpolation):
            return _tf.resize_image(image, size, interpolation=interpolation, max_size=max_size, antialias=antialias)
        return orig_image(image, size, interpolation=interpolation, max_size=max_size, antialias=antialias)

    def resize_video(video, size, interpolation=tv_geo.InterpolationMode.BILINEAR, max_size=None, antialias=True):
        return resize_image(video, size, interpolation=interpolation, max_size=max_size, antialias=antialias)

    registry[torch.Tensor] = resize_image
    registry[tv_tensors.Image] = tv_utils._kernel_tv_tensor_wrapper(resize_image)
    registry[tv_tensors.Video] = tv_utils._kernel_tv_tensor_wrapper(resize_video)

    _state.update(dict(tv_boxes=tv_boxes, torchvision=torchvision, orig_batched_nms=orig_batched_nms,
                       registry=registry, saved_registry=saved, tv_poolers=tv_poolers, orig_msra=orig_msra,
                       tv_roi_heads=tv_roi_heads, tv_rpn=tv_rpn, orig_pp=orig_pp, orig_fp=orig_fp,
                       tv_presets=tv_presets, orig_preset_forward=orig_preset_forward))


def uninstall() -> None:
    if not _state:
        return
    torch.ops.vision_b200._install(False)
    _state["tv_boxes"].batched_nms = _state["orig_batched_nms"]
    _state["torchvision"].ops.batched_nms = _state["orig_batched_nms"]
    _state["tv_poolers"]._multiscale_roi_align = _state["orig_msra"]
    _state["tv_presets"].ImageClassification.forward = _state["orig_preset_forward"]
    _state["tv_roi_heads"].RoIHeads.postprocess_detections = _state["orig_pp"]
    _state["tv_rpn"].RegionProposalNetwork.filter_proposals = _state["orig_fp"]
    reg = _state["registry"]
    reg.clear()
    reg.update(_state["saved_registry"])
    _state.clear()

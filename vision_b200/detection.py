"""Detection post-processing fused around batched_nms: drop-in bodies for
``RoIHeads.postprocess_detections`` (torchvision/models/detection/roi_heads.py:680-737) and
``RegionProposalNetwork.filter_proposals`` (rpn.py:242-298).  Everything up to the per-image loop (box decoding, softmax,
per-level top-k, sigmoid) is the reference's own tensor code; the per-image tail - clip, score filter, remove_small_boxes,
batched_nms, top-k, gathers - is ONE call of ``vision_b200::detection_postprocess`` per image."""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import Tensor

from . import _lib


def _fusable(t: Tensor) -> bool:
    return isinstance(t, Tensor) and t.is_cuda and t.dtype == torch.float32 and not torch.jit.is_scripting() and not torch.jit.is_tracing()


def detection_postprocess(boxes: Tensor, scores: Tensor, labels: Tensor, image_shape, score_thresh: float, score_inclusive: bool,
                          min_size: float, nms_thresh: float, topk: int):
    """clip_boxes_to_image -> score filter -> remove_small_boxes -> batched_nms -> keep[:topk] -> (boxes, scores, labels)."""
    _lib.load_ops()
    h, w = image_shape
    return torch.ops.vision_b200.detection_postprocess(boxes, scores, labels, float(h), float(w), float(score_thresh),
                                                       bool(score_inclusive), float(min_size), float(nms_thresh), int(topk))


def roi_heads_postprocess_detections(self, class_logits, box_regression, proposals, image_shapes, _orig=None):
    """RoIHeads.postprocess_detections with the per-image tail fused (same outputs, same order)."""
    if not _fusable(class_logits):
        return _orig(self, class_logits, box_regression, proposals, image_shapes)
    device = class_logits.device
    num_classes = class_logits.shape[-1]
    boxes_per_image = [boxes_in_image.shape[0] for boxes_in_image in proposals]
    pred_boxes = self.box_coder.decode(box_regression, proposals)
    pred_scores = F.softmax(class_logits, -1)
    pred_boxes_list = pred_boxes.split(boxes_per_image, 0)
    pred_scores_list = pred_scores.split(boxes_per_image, 0)
    all_boxes, all_scores, all_labels = [], [], []
    for boxes, scores, image_shape in zip(pred_boxes_list, pred_scores_list, image_shapes):
        n = scores.shape[0]
        labels = torch.arange(1, num_classes, device=device).view(1, -1).expand(n, num_classes - 1)
        # background column dropped, every class prediction becomes a separate instance (roi_heads.py:712-721)
        b, s, l = detection_postprocess(boxes[:, 1:].reshape(-1, 4), scores[:, 1:].reshape(-1), labels.reshape(-1), image_shape,
                                        self.score_thresh, False, 1e-2, self.nms_thresh, self.detections_per_img)
        all_boxes.append(b)
        all_scores.append(s)
        all_labels.append(l)
    return all_boxes, all_scores, all_labels


def rpn_filter_proposals(self, proposals, objectness, image_shapes, num_anchors_per_level, _orig=None):
    """RegionProposalNetwork.filter_proposals with the per-image tail fused."""
    if not _fusable(proposals):
        return _orig(self, proposals, objectness, image_shapes, num_anchors_per_level)
    num_images = proposals.shape[0]
    device = proposals.device
    objectness = objectness.detach().reshape(num_images, -1)
    levels = torch.cat([torch.full((n,), idx, dtype=torch.int64, device=device) for idx, n in enumerate(num_anchors_per_level)], 0)
    levels = levels.reshape(1, -1).expand_as(objectness)
    top_n_idx = self._get_top_n_idx(objectness, num_anchors_per_level)
    batch_idx = torch.arange(num_images, device=device)[:, None]
    objectness = objectness[batch_idx, top_n_idx]
    levels = levels[batch_idx, top_n_idx]
    proposals = proposals[batch_idx, top_n_idx]
    objectness_prob = torch.sigmoid(objectness)
    final_boxes, final_scores = [], []
    for boxes, scores, lvl, img_shape in zip(proposals, objectness_prob, levels, image_shapes):
        b, s, _ = detection_postprocess(boxes, scores, lvl, img_shape, self.score_thresh, True, self.min_size, self.nms_thresh,
                                        self.post_nms_top_n())
        final_boxes.append(b)
        final_scores.append(s)
    return final_boxes, final_scores

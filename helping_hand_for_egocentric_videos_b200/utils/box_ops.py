"""Mirror of the hot-path functions of the reference's ``utils/box_ops.py`` (:9-61) on the box kernels of
libhh_b200.so.  Same names, argument meaning, asserts and return values."""
from __future__ import annotations

import torch

from .. import ops


def box_cxcywh_to_xyxy(x):
    return ops.box_convert(x, to_xyxy=True)


def box_xyxy_to_cxcywh(x):
    return ops.box_convert(x, to_xyxy=False)


def box_iou(boxes1, boxes2):
    """(iou, union), iou = inter / (union + 1e-4) as in the reference (:36)."""
    iou, union, _ = ops.box_pairwise(boxes1, boxes2)
    return iou, union


def generalized_box_iou(boxes1, boxes2):
    """GIoU matrix [N,M] of xyxy boxes; degenerate boxes are rejected like the reference (:51-52)."""
    assert (boxes1[:, 2:] >= boxes1[:, :2]).all()
    assert (boxes2[:, 2:] >= boxes2[:, :2]).all()
    return ops.box_pairwise(boxes1, boxes2)[2]


def matcher_cost(pred_cxcywh, tgt_cxcywh, cost_bbox=5.0, cost_giou=2.0):
    """HungarianMatcher cost matrix with exclude_class=True (reference model/box_utils.py:75-88) in one kernel."""
    return ops.box_match_cost(pred_cxcywh, tgt_cxcywh, cost_bbox, cost_giou)

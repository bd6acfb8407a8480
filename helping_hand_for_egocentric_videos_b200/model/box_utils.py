"""Mirror of the training-side box utilities of the reference's ``model/box_utils.py``: HungarianMatcher (:15-92),
build_matcher (:95-96), SetCriterion (:99-238), prepare_targets (:249-279), split_detr_out (:433-443) and
compute_box_loss (:446-461).

Same constructor, ``forward(outputs, targets, exclude_class=False)`` signature and return value (a list with one
``(index_i, index_j)`` pair of int64 CPU tensors per image).  The reference copies the whole cost matrix to the host and
calls scipy once per image (:86-91); here the cost (hh_box_match_cost [+ hh_match_cost_class]) and every per-image
assignment (hh_assign) stay on the device and ONE small index array comes back.  With ``exclude_class=True`` (the only
way the reference calls it, :456) the 22 048-way softmax the reference computes and discards (:66) is not evaluated.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from .. import ops
from ..utils import box_ops


class HungarianMatcher(nn.Module):
    def __init__(self, cost_class: float = 1, cost_bbox: float = 1, cost_giou: float = 1):
        super().__init__()
        self.cost_class = cost_class
        self.cost_bbox = cost_bbox
        self.cost_giou = cost_giou
        assert cost_class != 0 or cost_bbox != 0 or cost_giou != 0, "all costs cant be 0"

    @torch.no_grad()
    def forward(self, outputs, targets, exclude_class=False):
        logits, boxes = outputs["pred_logits"], outputs["pred_boxes"]
        if not boxes.is_cuda:
            raise RuntimeError("HungarianMatcher (B200): pred_boxes is on %s; there is no CPU fallback" % boxes.device)
        bs, num_queries = logits.shape[:2]
        out_bbox = boxes.flatten(0, 1)
        sizes = getattr(targets, "counts", None) or [len(v["boxes"]) for v in targets]
        assert len(sizes) == bs
        M = sum(sizes)
        if M == 0:
            e = torch.empty(0, dtype=torch.int64)
            return [(e.clone(), e.clone()) for _ in range(bs)]
        tgt_bbox = _packed(targets, "boxes").to(out_bbox.device)
        C = ops.box_match_cost(out_bbox, tgt_bbox, float(self.cost_bbox), float(self.cost_giou))    # [bs*Q, M]
        if not exclude_class:
            tgt_ids = _packed(targets, "labels").to(out_bbox.device)
            ops.match_cost_class(C, logits.flatten(0, 1), tgt_ids, float(self.cost_class))
        starts = np.concatenate([[0], np.cumsum(sizes[:-1])]).astype(np.int64)
        offset = np.arange(bs, dtype=np.int64) * (num_queries * M) + starts
        ri, ci, cnt = ops.assign(C, offset, [M] * bs, [num_queries] * bs, sizes)
        K = ri.shape[1]
        packed = torch.cat([ri, ci, cnt.to(torch.int64)[:, None]], dim=1).cpu().numpy()               # one D2H
        counts = packed[:, 2 * K]
        if (counts < 0).any():
            raise ValueError("cost matrix is infeasible")            # scipy's message for inf / nan costs
        out = [(torch.from_numpy(packed[i, :k]), torch.from_numpy(packed[i, K:K + k])) for i, k in enumerate(counts)]
        # kept for SetCriterion.loss_boxes: lets it gather the matched pairs on the device without per-image indexing
        self._last = (out, ri, ci, counts, starts, num_queries)
        return out


def build_matcher(args):
    return HungarianMatcher(cost_class=1, cost_bbox=5, cost_giou=2)


# ---------------------------------------------------------------------------------------------- criterion
class _PackedTargets(list):
    """The per-image target dicts of prepare_targets plus the packed tensors they are views of, so the matcher and the
    criterion need not concatenate them again."""
    boxes = labels = counts = None


def _packed(targets, key):
    t = getattr(targets, key, None)
    return t if t is not None else torch.cat([v[key] for v in targets])


class _BoxLossFn(torch.autograd.Function):
    """(loss_bbox, loss_giou) of matched prediction / target pairs with the hand-written backward kernel."""

    @staticmethod
    def forward(ctx, pred_flat, src_row, tgt, num_boxes):
        from .. import _lib as L
        losses = torch.empty(2, dtype=torch.float32, device=pred_flat.device)
        K = src_row.numel()
        L.check(L.load().hh_box_loss_forward(L.ptr(pred_flat), L.ptr(src_row), L.ptr(tgt), K, float(num_boxes),
                                             L.ptr(losses), L.stream_ptr()), "hh_box_loss_forward")
        ctx.save_for_backward(pred_flat, src_row, tgt)
        ctx.num_boxes = float(num_boxes)
        return losses[0].clone(), losses[1].clone()

    @staticmethod
    def backward(ctx, g_bbox, g_giou):
        from .. import _lib as L
        pred_flat, src_row, tgt = ctx.saved_tensors
        g = torch.stack([g_bbox, g_giou]).float().contiguous()
        grad = torch.empty_like(pred_flat)
        L.check(L.load().hh_box_loss_backward(L.ptr(pred_flat), L.ptr(src_row), L.ptr(tgt), src_row.numel(), ctx.num_boxes,
                                              L.ptr(g), L.ptr(grad), pred_flat.shape[0], L.stream_ptr()),
                "hh_box_loss_backward")
        return grad, None, None, None


def is_dist_avail_and_initialized():
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized()


def get_world_size():
    import torch.distributed as dist
    return dist.get_world_size() if is_dist_avail_and_initialized() else 1


class SetCriterion(nn.Module):
    """Reference model/box_utils.py:99-238 (DETR criterion restricted, as there, to the 'boxes' and 'cardinality'
    losses): Hungarian matching of the last layer's outputs, then L1 + GIoU on the matched pairs.  Matching, both loss
    terms and their gradient are device kernels (hh_assign, hh_box_loss_*)."""

    def __init__(self, num_classes, matcher, weight_dict, eos_coef, losses):
        super().__init__()
        self.num_classes = num_classes
        self.matcher = matcher
        self.weight_dict = weight_dict
        self.eos_coef = eos_coef
        self.losses = losses
        empty_weight = torch.ones(num_classes + 1)
        empty_weight[-1] = self.eos_coef
        self.register_buffer('empty_weight', empty_weight)

    @torch.no_grad()
    def loss_cardinality(self, outputs, targets, indices, num_boxes, box_type):
        pred_logits = outputs['pred_logits']
        tgt_lengths = torch.as_tensor([len(v["labels"]) for v in targets], device=pred_logits.device)
        top = ops.row_argmax(pred_logits.flatten(0, 1)).view(pred_logits.shape[:2])
        card_pred = (top != pred_logits.shape[-1] - 1).sum(1)
        card_err = (card_pred.float() - tgt_lengths.float()).abs().mean()
        return {f'cardinality_error_{box_type}': card_err}

    def loss_boxes(self, outputs, targets, indices, num_boxes, box_type):
        assert 'pred_boxes' in outputs
        pred = outputs['pred_boxes']
        all_boxes = _packed(targets, "boxes")
        last = getattr(self.matcher, "_last", None)
        if last is not None and last[0] is indices:
            # indices straight from our matcher: pair lists are still on the device; only two small host index arrays
            # (image id and slot of every matched pair, known from the per-image counts) are uploaded
            _, ri, ci, counts, starts, nq = last
            bidx = np.repeat(np.arange(len(counts)), counts)
            kidx = np.arange(len(bidx)) - np.repeat(np.cumsum(counts) - counts, counts)
            sel = torch.from_numpy(np.stack([bidx, kidx, starts[bidx]])).to(pred.device, non_blocking=True)
            src_row = sel[0] * nq + ri[sel[0], sel[1]]
            target_boxes = all_boxes[sel[2] + ci[sel[0], sel[1]]]
        else:
            batch_idx, src_idx = self._get_src_permutation_idx(indices)
            src_row = (batch_idx * pred.shape[1] + src_idx).to(pred.device)
            starts_, flat = 0, []
            for t, (_, i) in zip(targets, indices):        # one gather over the packed targets (:164)
                flat.append(i + starts_)
                starts_ += len(t['boxes'])
            target_boxes = all_boxes[torch.cat(flat).to(all_boxes.device)]
        pred_flat = pred.flatten(0, 1)
        if pred_flat.dtype != torch.float32 or not pred_flat.is_contiguous():
            pred_flat = pred_flat.float().contiguous()
        loss_bbox, loss_giou = _BoxLossFn.apply(pred_flat, src_row.contiguous(),
                                                target_boxes.to(pred.device).float().contiguous(), num_boxes)
        return {f'loss_bbox_{box_type}': loss_bbox, f'loss_giou_{box_type}': loss_giou}

    def _get_src_permutation_idx(self, indices):
        batch_idx = torch.cat([torch.full_like(src, i) for i, (src, _) in enumerate(indices)])
        src_idx = torch.cat([src for (src, _) in indices])
        return batch_idx, src_idx

    def _get_tgt_permutation_idx(self, indices):
        batch_idx = torch.cat([torch.full_like(tgt, i) for i, (_, tgt) in enumerate(indices)])
        tgt_idx = torch.cat([tgt for (_, tgt) in indices])
        return batch_idx, tgt_idx

    def get_loss(self, loss, outputs, targets, indices, num_boxes, box_type, **kwargs):
        loss_map = {'cardinality': self.loss_cardinality, 'boxes': self.loss_boxes}
        assert loss in loss_map, f'do you really want to compute {loss} loss?'
        return loss_map[loss](outputs, targets, indices, num_boxes, box_type, **kwargs)

    def forward(self, outputs, targets, box_type, exclude_class=False):
        outputs_without_aux = {k: v for k, v in outputs.items() if k != 'aux_outputs'}
        indices_last = self.matcher(outputs_without_aux, targets, exclude_class=exclude_class)
        num_boxes = sum(len(t["labels"]) for t in targets)
        if is_dist_avail_and_initialized():
            nb = torch.as_tensor([num_boxes], dtype=torch.float, device=next(iter(outputs.values())).device)
            torch.distributed.all_reduce(nb)
            num_boxes = nb.item()
        num_boxes = max(float(num_boxes) / get_world_size(), 1.0)
        losses = {}
        for loss in self.losses:
            losses.update(self.get_loss(loss, outputs, targets, indices_last, num_boxes, box_type))
        if 'aux_outputs' in outputs:
            for i, aux_outputs in enumerate(outputs['aux_outputs']):
                indices = self.matcher(aux_outputs, targets, exclude_class=exclude_class)
                for loss in self.losses:
                    l_dict = self.get_loss(loss, aux_outputs, targets, indices, num_boxes, box_type)
                    losses.update({k + f'_{i}': v for k, v in l_dict.items()})
        return losses, indices_last


@torch.no_grad()
def prepare_targets(boxes, classes, image_size, center_crop=True):
    """Reference :249-279: xyxy pixel boxes -> per-image dicts of normalised cxcywh boxes (absent / degenerate boxes
    dropped).  Unlike the reference (:255 `.cuda()`), tensors stay on the device of `boxes`."""
    if classes is None:
        classes = 1 - (boxes.sum(-1) != 0).float()            # dummy labels, as the reference builds them (:254)
    if center_crop:
        shift = torch.zeros_like(boxes)
        dis = (image_size[:, 1] - image_size[:, 0]) / 2
        wide, tall = dis >= 0, dis < 0
        shift[wide, :, 0] = -dis[wide, None]
        shift[wide, :, 2] = -dis[wide, None]
        shift[tall, :, 1] = dis[tall, None]
        shift[tall, :, 3] = dis[tall, None]
        boxes += shift
        boxes = torch.clip(boxes, min=0, max=256).div(256)
    else:
        boxes = torch.clip(boxes, min=0, max=224).div(224)
    # all images at once (the reference loops over images with one boolean index each, :270-278): one mask, one box
    # conversion, one host read of the per-image counts; the per-image dicts are views into the packed tensors
    avail = (classes != -1) & (boxes[..., 2] > boxes[..., 0]) & (boxes[..., 3] > boxes[..., 1])
    counts = avail.sum(1).tolist()
    packed_boxes = box_ops.box_xyxy_to_cxcywh(boxes[avail])
    packed_labels = classes[avail]
    out = _PackedTargets({'labels': l, 'boxes': b} for l, b in zip(packed_labels.split(counts), packed_boxes.split(counts)))
    out.boxes, out.labels, out.counts = packed_boxes, packed_labels, counts
    return out


def split_detr_out(detr_out, start=0, end=2):
    """Reference :433-443, including its quirk: 'aux_outputs' is emptied before it is iterated, so the auxiliary
    decoder layers never contribute a box loss."""
    out = detr_out.copy()
    out['pred_boxes'] = detr_out['pred_boxes'][:, start:end, :]
    out['pred_logits'] = detr_out['pred_logits'][:, start:end]
    out['aux_outputs'] = []
    return out


def compute_box_loss(box_type, criterion, detr_out, target_boxes, target_classes, all_image_size, n_queries=10):
    """Reference :446-461."""
    targets = prepare_targets(target_boxes, target_classes, all_image_size, center_crop=False)
    if box_type == 'hand_boxes':
        detr_pred = split_detr_out(detr_out, start=0, end=2)
    elif box_type == 'obj_boxes':
        detr_pred = split_detr_out(detr_out, start=2, end=n_queries)
    elif box_type == 'all_boxes':
        detr_pred = detr_out
    detr_loss_dict, matched_indices = criterion(detr_pred, targets, box_type, exclude_class=True)
    weight_dict = criterion.weight_dict
    for k in detr_loss_dict.keys():
        if k in weight_dict:
            detr_loss_dict[k] *= weight_dict[k]
    return sum(v for k, v in detr_loss_dict.items() if k in weight_dict) / (len(weight_dict) / 3), matched_indices

"""Mirror of the matcher of the reference's ``model/box_utils.py`` (HungarianMatcher :15-92, build_matcher :95-96).

Same constructor, ``forward(outputs, targets, exclude_class=False)`` signature and return value (a list with one
``(index_i, index_j)`` pair of int64 CPU tensors per image).  The reference copies the whole cost matrix to the host and
calls scipy once per image (:86-91); here the cost (hh_box_match_cost [+ hh_match_cost_class]) and every per-image
assignment (hh_assign) stay on the device and ONE small index array comes back.  With ``exclude_class=True`` (the only
way the reference calls it, :456) the 22 048-way softmax the reference computes and discards (:66) is not evaluated.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import ops


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
        sizes = [len(v["boxes"]) for v in targets]
        assert len(sizes) == bs
        M = sum(sizes)
        if M == 0:
            e = torch.empty(0, dtype=torch.int64)
            return [(e.clone(), e.clone()) for _ in range(bs)]
        tgt_bbox = torch.cat([v["boxes"] for v in targets]).to(out_bbox.device)
        C = ops.box_match_cost(out_bbox, tgt_bbox, float(self.cost_bbox), float(self.cost_giou))    # [bs*Q, M]
        if not exclude_class:
            tgt_ids = torch.cat([v["labels"] for v in targets]).to(out_bbox.device)
            ops.match_cost_class(C, logits.flatten(0, 1), tgt_ids, float(self.cost_class))
        starts = [0]
        for s in sizes[:-1]:
            starts.append(starts[-1] + s)
        offset = [i * num_queries * M + starts[i] for i in range(bs)]
        ri, ci, cnt = ops.assign(C, offset, [M] * bs, [num_queries] * bs, sizes)
        K = ri.shape[1]
        packed = torch.cat([ri, ci, cnt.to(torch.int64)[:, None]], dim=1).cpu()                      # one D2H
        out = []
        for i in range(bs):
            k = int(packed[i, 2 * K])
            if k < 0:
                raise ValueError("cost matrix is infeasible")        # scipy's message for inf / nan costs
            out.append((packed[i, :k].clone(), packed[i, K:K + k].clone()))
        return out


def build_matcher(args):
    return HungarianMatcher(cost_class=1, cost_bbox=5, cost_giou=2)

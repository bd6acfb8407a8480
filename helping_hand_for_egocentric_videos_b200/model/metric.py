"""Mirror of the hot-path functions of the reference's ``model/metric.py``: ``sim_matrix`` (:363-375) and
``egomcq_accuracy_metrics`` (:209-225), evaluated by the fused scoring kernels of libhh_b200.so."""
from __future__ import annotations

import torch

from .. import ops


class _SimMatrixFn(torch.autograd.Function):
    """sim_matrix with its backward kernel (hh_sim_matrix_backward), for the training-side callers
    (reference run/train.py:136, model/loss.py:96)."""

    @staticmethod
    def forward(ctx, a, b, eps):
        ctx.save_for_backward(a, b)
        ctx.eps = eps
        return ops.sim_matrix(a, b, eps)

    @staticmethod
    def backward(ctx, grad):
        a, b = ctx.saved_tensors
        da, db = ops.sim_matrix_backward(a, b, grad.contiguous(), ctx.eps, ctx.needs_input_grad[0],
                                         ctx.needs_input_grad[1])
        return da, db, None


def sim_matrix(a, b, eps=1e-8, norm=True):
    """L2-normalise (norms clamped at eps) + similarity in one kernel.  2-D inputs -> mm, 3-D -> bmm, as the reference."""
    if not norm:
        raise NotImplementedError("sim_matrix(norm=False) is never used by the reference scripts")
    if a.dim() == 2:
        if torch.is_grad_enabled() and (a.requires_grad or b.requires_grad):
            return _SimMatrixFn.apply(a.float().contiguous(), b.float().contiguous(), eps)
        return ops.sim_matrix(a, b, eps)
    if a.dim() == 3:       # bmm form: one scoring launch per batch entry, each carrying its own backward
        return torch.stack([sim_matrix(x, y, eps) for x, y in zip(a, b)])
    raise ValueError("sim_matrix expects 2-D or 3-D inputs")


def egomcq_choices(preds) -> torch.Tensor:
    """argmax over the 5 options of every question: preds [G,1,5] (or [G,5]) -> int64 [G]."""
    p = preds.reshape(preds.shape[0], -1)
    return ops.row_argmax(p)


def egomcq_accuracy_metrics(preds, labels, types):
    """Same result dict as the reference (:209-225): accuracy in percent per question type; the first sorted type is
    reported as "Intra-video", the second as "Inter-video".  One argmax kernel for all questions instead of a Python
    loop with .item() per question."""
    metrics = {}
    preds = preds if preds.is_cuda else preds.cuda()
    choice = egomcq_choices(preds).cpu()
    labels = labels.reshape(-1).cpu()
    types = types.reshape(-1).cpu()
    for type_i, group_i in zip(torch.unique(types), ["Intra-video", "Inter-video"]):
        sel = types == type_i
        correct = int((choice[sel] == labels[sel]).sum())
        metrics[group_i] = correct / int(sel.sum()) * 100
    return metrics

"""Mirror of the reference's ``model/loss.py``: ``EgoNCE`` (:8-70) and ``WordContrastiveLoss`` (:72-106) with the same
constructors, ``forward`` signatures and return values, evaluated (forward AND backward) by the loss kernels of
libhh_b200.so through ``torch.autograd.Function`` shims.  No host synchronisation: the reference's per-clip scipy
assignment loop (:88-93) is one hh_assign launch inside hh_word_loss_forward.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import _lib as L


def _f32c(t):
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()


class _EgoNCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mask_v, mask_n, pad, R, temperature, vn_threshold):
        N, M = x.shape
        dev = x.device
        mask_bool = torch.empty(N, M, dtype=torch.uint8, device=dev)
        keep = torch.empty(N, dtype=torch.uint8, device=dev)
        saved = torch.empty(3 * N + 3 * M + 2, dtype=torch.float32, device=dev)
        L.check(L.load().hh_egonce_forward(L.ptr(x), N, M, L.ptr(mask_v), L.ptr(mask_n), R, L.ptr(pad), temperature,
                                           float(vn_threshold), L.ptr(mask_bool), L.ptr(keep), L.ptr(saved),
                                           L.stream_ptr()), "hh_egonce_forward")
        ctx.save_for_backward(x, mask_bool, keep, saved)
        ctx.temperature = temperature
        ctx.mark_non_differentiable(mask_bool, keep)
        return saved[3 * N + 3 * M].clone(), mask_bool, keep

    @staticmethod
    def backward(ctx, g, _g_mask, _g_keep):
        x, mask_bool, keep, saved = ctx.saved_tensors
        N, M = x.shape
        g = _f32c(g.reshape(1))
        gx = torch.empty_like(x)
        L.check(L.load().hh_egonce_backward(L.ptr(x), N, M, ctx.temperature, L.ptr(mask_bool), L.ptr(keep), L.ptr(saved),
                                            L.ptr(g), L.ptr(gx), L.stream_ptr()), "hh_egonce_backward")
        return gx, None, None, None, None, None, None


class EgoNCE(nn.Module):
    def __init__(self, temperature=0.07, noun=True, verb=True):
        super().__init__()
        self.noun = noun
        self.verb = verb
        self.temperature = temperature

    def forward(self, x, mask_v, mask_n, multi_pad_mask=None, strict_mask=False, vn_threshold=0):
        """x [N, M] similarity matrix.  Returns (loss, mask_bool) like the reference (:15-70): mask_bool is the boolean
        positive mask of the rows that survive the pad filter."""
        if not x.is_cuda:
            raise RuntimeError("EgoNCE (B200): x is on %s; there is no CPU fallback" % x.device)
        if mask_v is None and mask_n is None:
            raise UnboundLocalError("EgoNCE needs mask_v and/or mask_n (the reference leaves `mask` undefined otherwise)")
        N, M = x.shape
        if multi_pad_mask is None:
            if N != M:
                raise RuntimeError("single-positive EgoNCE needs a square similarity matrix, got %s" % (tuple(x.shape),))
            R, pad = 1, None
        else:
            if tuple(multi_pad_mask.shape) != (N, M):
                raise RuntimeError("multi_pad_mask must have the shape of x")
            R = N // M
            if R * M != N:
                raise RuntimeError("rows of x must be a multiple of its columns (R captions per video)")
            if (mask_v is None or mask_n is None) and R != 5:
                raise RuntimeError("the single-mask multi-positive branches of the reference repeat the mask 5 times "
                                   "(model/loss.py:49,54); got %d captions per video" % R)
            pad = _f32c(multi_pad_mask)
        for m in (mask_v, mask_n):
            if m is not None and tuple(m.shape) != (M, M):
                raise RuntimeError("mask_v / mask_n must be [%d, %d]" % (M, M))
        loss, mask_bool, keep = _EgoNCEFn.apply(_f32c(x), None if mask_v is None else _f32c(mask_v),
                                                None if mask_n is None else _f32c(mask_n), pad, R,
                                                float(self.temperature), float(vn_threshold))
        mask_bool = mask_bool.bool()
        if multi_pad_mask is not None:
            mask_bool = mask_bool[keep.bool()]
        return loss, mask_bool


class _WordLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, nouns, pred, inds, temperature, noun_threshold):
        V, d = nouns.shape
        B2, Q, _ = pred.shape
        Wm = inds.shape[1]
        dev = nouns.device
        lib = L.load()
        S = B2 * Wm
        col_ind = torch.empty(B2, Wm, dtype=torch.int64, device=dev)
        sel = torch.empty(S, d, dtype=torch.float32, device=dev)
        sel_row = torch.empty(S, dtype=torch.int64, device=dev)
        dlogits = torch.empty(S, V, dtype=torch.float32, device=dev)
        stats = torch.zeros(4, dtype=torch.float32, device=dev)
        ws = torch.empty(lib.hh_word_loss_workspace_bytes(V, d, B2, Q, Wm), dtype=torch.uint8, device=dev)
        L.check(lib.hh_word_loss_forward(L.ptr(nouns), V, d, L.ptr(pred), B2, Q, L.ptr(inds), Wm, temperature,
                                         noun_threshold, L.ptr(col_ind), L.ptr(sel), L.ptr(sel_row), L.ptr(dlogits),
                                         L.ptr(stats), L.ptr(ws), L.stream_ptr()), "hh_word_loss_forward")
        ctx.save_for_backward(nouns, sel, sel_row, dlogits, stats)
        ctx.dims = (V, d, B2, Q, Wm)
        ctx.mark_non_differentiable(col_ind)
        return stats[0].clone(), col_ind

    @staticmethod
    def backward(ctx, g, _g_col):
        nouns, sel, sel_row, dlogits, stats = ctx.saved_tensors
        V, d, B2, Q, Wm = ctx.dims
        lib = L.load()
        g = _f32c(g.reshape(1))
        need_n, need_p = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        d_nouns = torch.empty(V, d, dtype=torch.float32, device=nouns.device) if need_n else None
        d_pred = torch.empty(B2, Q, d, dtype=torch.float32, device=nouns.device) if need_p else None
        ws = torch.empty(lib.hh_word_loss_workspace_bytes(V, d, B2, Q, Wm), dtype=torch.uint8, device=nouns.device)
        L.check(lib.hh_word_loss_backward(L.ptr(nouns), V, d, B2, Q, Wm, L.ptr(sel), L.ptr(sel_row), L.ptr(dlogits),
                                          L.ptr(stats), L.ptr(g), L.ptr(d_pred), L.ptr(d_nouns), L.ptr(ws),
                                          L.stream_ptr()), "hh_word_loss_backward")
        return d_nouns, d_pred, None, None, None


class WordContrastiveLoss(nn.Module):
    def __init__(self, temperature=0.07, noun_threshold=0.6):
        super().__init__()
        self.temperature = temperature
        self.noun_threshold = noun_threshold
        self.last_col_ind = None      # int64 [B, W]: query matched to each ground-truth noun slot (-1 = empty slot)

    def forward(self, noun_embeds, pred_noun_embeds, noun_gt_inds):
        """noun_embeds [V, d]; pred_noun_embeds [B, Q, d]; noun_gt_inds int64 [B, W] (0 = no noun) -> scalar loss.
        A batch without any noun gives NaN (the reference raises from torch.cat of an empty list, :94)."""
        if not noun_embeds.is_cuda:
            raise RuntimeError("WordContrastiveLoss (B200): inputs are on %s; there is no CPU fallback"
                               % noun_embeds.device)
        loss, col_ind = _WordLossFn.apply(_f32c(noun_embeds), _f32c(pred_noun_embeds),
                                          noun_gt_inds.to(torch.int64).contiguous(), float(self.temperature),
                                          float(self.noun_threshold))
        self.last_col_ind = col_ind
        return loss

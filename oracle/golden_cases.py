"""Golden-vector case table shared by `make_golden.py` (reference side) and the tests (oracle / CUDA side).

TEST INFRASTRUCTURE ONLY.  Inputs and weights are regenerated from seeds; fixtures store outputs only.
"""
import os

import torch

from . import hh_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

_TXT = dict(text_width=768, text_heads=12, text_layers=1, vocab=128)

CASES = {
    # tiny encoder, every quirk exercised: n=16 patches, T=3 frames (odd), 2 heads of 64
    "enc_tiny": dict(kind="encoder", seed=11, B=2, G=3,
                     cfg=dict(img=56, patch=14, D=128, L=2, H=2, T=3, **_TXT), stride=(1, 1)),
    # depth-1 truncation of the same shapes: pins block 0 in isolation
    "enc_tiny_d1": dict(kind="encoder", seed=12, B=1, G=1,
                        cfg=dict(img=56, patch=14, D=128, L=1, H=2, T=3, **_TXT), stride=(1, 1)),
    # BASELINE config c0 geometry (TimeSformer-B/16, 4 frames, batch 2); outputs subsampled to stay small
    "enc_c0": dict(kind="encoder", seed=13, B=2, G=2,
                   cfg=dict(img=224, patch=16, D=768, L=12, H=12, T=4, **_TXT), stride=(29, 7)),
    "dec_tiny_traj": dict(kind="decoder", seed=21, B=2, G=3,
                          cfg=dict(C=128, heads=2, layers=2, ffn=256, Q=5, n=16, T=3, F=128, ncls=30, pred_traj=True)),
    "dec_tiny_notraj": dict(kind="decoder", seed=22, B=3, G=2,
                            cfg=dict(C=128, heads=2, layers=3, ffn=256, Q=13, n=16, T=2, F=192, ncls=30,
                                     pred_traj=False)),
    # c0 decoder geometry: C=512, 8 heads, 6 layers, Q=5 (nq=4), n=196, T=4, F=768, full 22048-way class head
    "dec_c0": dict(kind="decoder", seed=23, B=2, G=2,
                   cfg=dict(C=512, heads=8, layers=6, ffn=2048, Q=5, n=196, T=4, F=768, ncls=22047, pred_traj=True),
                   logit_stride=37),
    "boxes": dict(kind="boxes", seed=31, N=40, M=7),
    "score": dict(kind="score", seed=41, Na=6, Nb=10, d=256, G=12),
}


def backbone_state_dict(case):
    c = case["cfg"]
    n = (c["img"] // c["patch"]) ** 2
    shapes = O.clip_param_shapes(c["D"], c["L"], c["patch"], n, c["T"], text_width=c["text_width"],
                                 text_layers=c["text_layers"], vocab=c["vocab"])
    return O.synth_state_dict(shapes, case["seed"])


def decoder_state_dict(case):
    c = case["cfg"]
    shapes = O.decoder_param_shapes(c["C"], c["Q"], c["n"], c["T"], c["F"], c["ncls"] + 1, layers=c["layers"],
                                    ffn=c["ffn"], pred_traj=c["pred_traj"])
    return O.synth_state_dict(shapes, case["seed"])


def make_tokens(G, vocab, g, ctx=77):
    """Random captions: ids in [1, vocab-2], EOT = vocab-1 (row max, picked by argmax, LaviLa.py:669), zero padded."""
    tok = torch.zeros(G, ctx, dtype=torch.long)
    for i in range(G):
        ln = int(torch.randint(3, ctx - 1, (1,), generator=g))
        tok[i, :ln] = torch.randint(1, vocab - 1, (ln,), generator=g)
        tok[i, ln] = vocab - 1
    return tok


def make_inputs(case):
    g = torch.Generator().manual_seed(case["seed"] + 1000)
    kind = case["kind"]
    if kind == "encoder":
        c = case["cfg"]
        video = torch.randn(case["B"], c["T"], 3, c["img"], c["img"], generator=g)
        return video, make_tokens(case["G"], c["vocab"], g)
    if kind == "decoder":
        c = case["cfg"]
        feats = torch.randn(case["B"], c["T"], c["n"], c["F"], generator=g)
        return feats, torch.randn(case["G"], 768, generator=g)
    if kind == "boxes":
        def rb(k):
            cxy = 0.2 + 0.6 * torch.rand(k, 2, generator=g)
            wh = 0.02 + 0.35 * torch.rand(k, 2, generator=g)
            return torch.cat([cxy, wh], -1)
        p, t = rb(case["N"]), rb(case["M"])
        t[0] = p[3]            # identical pair: GIoU != 1 because of the +1e-4 (utils/box_ops.py:36)
        t[1, 2:] = 0.0         # degenerate (zero-area) target
        return p, t
    if kind == "score":
        a = torch.randn(case["Na"], case["d"], generator=g)
        b = torch.randn(case["Nb"], case["d"], generator=g)
        a[1] = 0.0             # zero row: exercises the eps clamp (model/metric.py:368-370)
        G = case["G"]
        preds = torch.randn(G, 1, 5, generator=g)
        labels = torch.randint(0, 5, (G,), generator=g)
        types = torch.randint(1, 3, (G,), generator=g)
        return a, b, preds, labels, types
    raise ValueError(kind)


def subsample(case, res):
    """Keep fixtures small: big tensors are strided (same rule applied to the oracle/CUDA side by the tests)."""
    out = {}
    for k, v in res.items():
        if k in ("image_feature_map", "text_feature_map") and "stride" in case:
            s0, s1 = case["stride"]
            v = v[:, ::s0, ::s1]
        if k in ("pred_logits", "aux_logits0") and "logit_stride" in case:
            v = v[..., ::case["logit_stride"]]
        out[k] = v.contiguous().clone() if isinstance(v, torch.Tensor) else v
    return out


def run_oracle(case):
    """The restatement's answer for a case, in the same (subsampled) form as the fixture."""
    kind = case["kind"]
    with torch.no_grad():
        if kind == "encoder":
            c = case["cfg"]
            sd = backbone_state_dict(case)
            video, tokens = make_inputs(case)
            out = O.clip_forward(video, tokens, sd, heads=c["H"], text_heads=c["text_heads"])
            return subsample(case, {k: out[k] for k in ("image_embed", "text_embed", "image_feature_map",
                                                        "text_feature_map")})
        if kind == "decoder":
            c = case["cfg"]
            sd = decoder_state_dict(case)
            feats, text_feat = make_inputs(case)
            out, hs, _, _ = O.decoder_forward(feats, sd, heads=c["heads"], pred_traj=c["pred_traj"])
            vid = O.obj_proj(hs[-1], sd)[:, -1]
            txt = O.txt_proj(text_feat, sd)
            res = {"pred_logits": out["pred_logits"], "pred_boxes": out["pred_boxes"], "hs": hs,
                   "aux_boxes": torch.stack([a["pred_boxes"] for a in out["aux_outputs"]]),
                   "aux_logits0": out["aux_outputs"][0]["pred_logits"],
                   "video_embed": vid, "text_embed": txt, "sim": O.sim_matrix(txt, vid)}
            return subsample(case, res)
        if kind == "boxes":
            p, t = make_inputs(case)
            pxy, txy = O.box_cxcywh_to_xyxy(p), O.box_cxcywh_to_xyxy(t)
            iou, union = O.box_iou(pxy, txy)
            return {"xyxy": pxy, "back": O.box_xyxy_to_cxcywh(pxy), "iou": iou, "union": union,
                    "giou": O.generalized_box_iou(pxy, txy), "cost": O.matcher_cost(p, t)}
        if kind == "score":
            a, b, preds, labels, types = make_inputs(case)
            return {"sim": O.sim_matrix(a, b), "sim3": O.sim_matrix(a[None], b[None]),
                    "acc": O.egomcq_accuracy(preds, labels, types)}
    raise ValueError(kind)

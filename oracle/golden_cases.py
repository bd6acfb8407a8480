"""Golden-vector case table shared by `make_golden.py` (reference side) and the tests (oracle / CUDA side).

TEST INFRASTRUCTURE ONLY.  Inputs and weights are regenerated from seeds; fixtures store outputs only.
"""
import os

import numpy as np
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
    # BASELINE c1 / c2 encoder geometry (TimeSformer-L/14: D=1024, 24 layers, 16 heads, 256 patches), 4 frames, 1 clip
    "enc_l14": dict(kind="encoder", seed=14, B=1, G=2,
                    cfg=dict(img=224, patch=14, D=1024, L=24, H=16, T=4, **_TXT), stride=(41, 13)),
    # the headline (c2 / c3) encoder geometry: the same tower at 16 frames (N = 4097 tokens), 1 clip
    "enc_l14_t16": dict(kind="encoder", seed=16, B=1, G=2,
                        cfg=dict(img=224, patch=14, D=1024, L=24, H=16, T=16, **_TXT), stride=(163, 13)),
    # full LARGE text tower (12 layers x 768, 12 heads, vocabulary 49408; model/LaviLa.py:141-147) beside a tiny visual tower
    "txt_large": dict(kind="encoder", seed=15, B=1, G=6,
                      cfg=dict(img=56, patch=14, D=128, L=1, H=2, T=3, text_width=768, text_heads=12, text_layers=12,
                               vocab=49408), stride=(7, 5)),
    "dec_tiny_traj": dict(kind="decoder", seed=21, B=2, G=3,
                          cfg=dict(C=128, heads=2, layers=2, ffn=256, Q=5, n=16, T=3, F=128, ncls=30, pred_traj=True)),
    "dec_tiny_notraj": dict(kind="decoder", seed=22, B=3, G=2,
                            cfg=dict(C=128, heads=2, layers=3, ffn=256, Q=13, n=16, T=2, F=192, ncls=30,
                                     pred_traj=False)),
    # c0 decoder geometry: C=512, 8 heads, 6 layers, Q=5 (nq=4), n=196, T=4, F=768, full 22048-way class head
    "dec_c0": dict(kind="decoder", seed=23, B=2, G=2,
                   cfg=dict(C=512, heads=8, layers=6, ffn=2048, Q=5, n=196, T=4, F=768, ncls=22047, pred_traj=True),
                   logit_stride=37),
    # BASELINE c2 decoder geometry (the headline configuration): nq = 12 -> Q = 13, 16 frames x 256 patches of 1024-d
    # features, no trajectory head (run/test_epic.py:150-153), full class head
    "dec_c2": dict(kind="decoder", seed=25, B=1, G=2,
                   cfg=dict(C=512, heads=8, layers=6, ffn=2048, Q=13, n=256, T=16, F=1024, ncls=22047, pred_traj=False),
                   logit_stride=37),
    # BASELINE c1 decoder geometry (run/test_EgoMCQ.py:236-246): nq = 4 -> Q = 5, 4 frames x 256 patches of 1024-d features,
    # trajectory head, full class head
    "dec_c1": dict(kind="decoder", seed=26, B=1, G=2,
                   cfg=dict(C=512, heads=8, layers=6, ffn=2048, Q=5, n=256, T=4, F=1024, ncls=22047, pred_traj=True),
                   logit_stride=37),
    # BASELINE c4 decoder geometry (run/train.py:448-458): nq = 12 -> Q = 13 at the same 4 x 256 x 1024 memory, trajectory head
    "dec_c4": dict(kind="decoder", seed=27, B=2, G=2,
                   cfg=dict(C=512, heads=8, layers=6, ffn=2048, Q=13, n=256, T=4, F=1024, ncls=22047, pred_traj=True),
                   logit_stride=37),
    # the decoder in train() mode: dropout at six sites per layer (nn.Dropout x4, both nn.MultiheadAttention modules'
    # attention probabilities) with the masks of hh_oracle.philox_keep injected into the reference through
    # torch.nn.functional.dropout; outputs and the reference's autograd gradients of a fixed linear functional
    "dec_train_dropout": dict(kind="decoder_train", seed=24, B=2, G=3,
                              cfg=dict(C=128, heads=2, layers=2, ffn=256, Q=5, n=16, T=3, F=128, ncls=30, pred_traj=True),
                              dropout=dict(p=0.1, seed=0x5EED0123456789, offset=7)),
    "boxes": dict(kind="boxes", seed=31, N=40, M=7),
    "score": dict(kind="score", seed=41, Na=6, Nb=10, d=256, G=12),
    # training-side rows a17-a19: matcher indices, EgoNCE (multi-positive, single-positive, noun-only), word loss,
    # with the reference's own autograd gradients
    "losses": dict(kind="losses", seed=51, Nv=6, R=5, d=32, V=50, B2=6, Q=12, Wm=4, bs=5, nq=10, ncls=20,
                   sizes=(3, 0, 4, 1, 2)),
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
    if kind in ("decoder", "decoder_train"):
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
    if kind == "losses":
        c = case
        Nv, R, d = c["Nv"], c["R"], c["d"]
        vid = torch.randn(Nv, d, generator=g)
        txt = vid.repeat_interleave(R, 0) * 0.6 + torch.randn(Nv * R, d, generator=g)
        verb = (torch.rand(Nv, 12, generator=g) < 0.25).float()
        noun = (torch.rand(Nv, 20, generator=g) < 0.2).float()
        verb[1] = verb[0]
        noun[1] = noun[0]                      # videos 0/1 share verbs and nouns: extra positives
        pad = (torch.rand(Nv * R, generator=g) > 0.3).float()
        pad[::R] = 1.0                          # the original caption of every clip is never padding
        pad = pad[:, None].repeat(1, Nv)
        nouns = torch.randn(c["V"], d, generator=g)
        nouns[7] = nouns[3] + 0.05 * torch.randn(d, generator=g)      # near-synonyms (cos > 0.6): masked logits
        nouns[11] = nouns[3] + 0.3 * torch.randn(d, generator=g)
        pred = torch.randn(c["B2"], c["Q"], d, generator=g)
        inds = torch.randint(1, c["V"], (c["B2"], c["Wm"]), generator=g)
        inds[0] = torch.tensor([3, 0, 7, 0])    # holes in the middle, synonyms as targets
        inds[1] = 0                             # clip without nouns
        inds[2, 1:] = 0
        inds[3, 0] = 0
        inds[4] = torch.tensor([11, 11, 3, 5])  # repeated noun
        pred[4, 2] = pred[4, 9]                 # identical queries: tied assignment costs

        def rb(k):
            return torch.cat([0.2 + 0.6 * torch.rand(k, 2, generator=g), 0.02 + 0.35 * torch.rand(k, 2, generator=g)], -1)
        pb = rb(c["bs"] * c["nq"]).view(c["bs"], c["nq"], 4)
        pb[2, 4] = pb[2, 1]                     # duplicate predictions: tied matcher costs
        pl = torch.randn(c["bs"], c["nq"], c["ncls"], generator=g)
        tb = [rb(k) for k in c["sizes"]]
        tl = [torch.randint(0, c["ncls"], (k,), generator=g) for k in c["sizes"]]
        # compute_box_loss inputs: 8 frames x 12 queries (2 hand + 10 object), pixel xyxy targets [frames, 4, 4] with
        # absent (all-zero) and degenerate rows; drawn last so the tensors above keep their values
        det_boxes = rb(8 * 12).view(8, 12, 4)
        det_logits = torch.randn(8, 12, c["ncls"], generator=g)
        lo = 224 * torch.rand(8, 4, 2, generator=g) * 0.7
        px = torch.cat([lo, lo + 10 + 60 * torch.rand(8, 4, 2, generator=g)], -1)
        px[torch.rand(8, 4, generator=g) < 0.3] = 0.0
        px[1, 0] = torch.tensor([50., 60., 50., 90.])        # zero width: dropped by prepare_targets
        px[2, :2] = 0.0                                        # a frame without hands
        return dict(vid=vid, txt=txt, verb=verb, noun=noun, pad=pad, nouns=nouns, pred=pred, inds=inds, pred_boxes=pb,
                    pred_logits=pl, tgt_boxes=tb, tgt_labels=tl, det_boxes=det_boxes, det_logits=det_logits, det_px=px)
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


def run_losses(inp, sim_matrix, egonce, word_loss, matcher, box_loss):
    """Shared driver of the 'losses' case: the same sequence of calls is made with the reference's modules
    (make_golden.py), the oracle's restatements (run_oracle) and the CUDA mirrors (tests/test_gpu_losses.py).
    egonce(x, mask_v, mask_n, pad) -> (loss, mask_bool); word_loss(nouns, pred, inds) -> (loss, flat col indices);
    matcher(outputs, targets, exclude_class) -> [(i, j)]; box_loss(detr_out, pixel_boxes, box_type) -> (loss, [(i, j)])
    = compute_box_loss(box_type, criterion, detr_out, boxes, None, sizes, n_queries=12)."""
    res = {}
    vid = inp["vid"].clone().requires_grad_(True)
    txt = inp["txt"].clone().requires_grad_(True)
    sim_v, sim_n = sim_matrix(inp["verb"], inp["verb"]), sim_matrix(inp["noun"], inp["noun"])
    x = sim_matrix(txt, vid)
    res["x"] = x.detach()
    loss, mb = egonce(x, sim_v, sim_n, inp["pad"])
    loss.backward()
    res.update(nce_multi=loss.detach(), nce_multi_mask=mb, nce_multi_dvid=vid.grad.clone(), nce_multi_dtxt=txt.grad.clone())
    R = inp["txt"].shape[0] // inp["vid"].shape[0]
    t1 = inp["txt"][::R].clone().requires_grad_(True)
    x1 = sim_matrix(t1, inp["vid"])
    loss, mb = egonce(x1, sim_v, sim_n, None)
    loss.backward()
    res.update(nce_single=loss.detach(), nce_single_mask=mb, nce_single_dtxt=t1.grad.clone())
    loss, mb = egonce(x1.detach(), None, sim_n, None)
    res.update(nce_noun=loss.detach(), nce_noun_mask=mb)
    nouns = inp["nouns"].clone().requires_grad_(True)
    pred = inp["pred"].clone().requires_grad_(True)
    loss, cols = word_loss(nouns, pred, inp["inds"])
    loss.backward()
    res.update(word=loss.detach(), word_cols=cols, word_dnouns=nouns.grad.clone(), word_dpred=pred.grad.clone())
    outputs = {"pred_logits": inp["pred_logits"], "pred_boxes": inp["pred_boxes"]}
    targets = [{"boxes": b, "labels": l} for b, l in zip(inp["tgt_boxes"], inp["tgt_labels"])]
    for name, excl in (("match_excl", True), ("match_cls", False)):
        idx = matcher(outputs, targets, excl)
        res[name + "_i"] = torch.cat([i for i, _ in idx])
        res[name + "_j"] = torch.cat([j for _, j in idx])
        res[name + "_n"] = torch.tensor([len(i) for i, _ in idx])
    det = inp["det_boxes"].clone().requires_grad_(True)
    detr_out = {"pred_boxes": det, "pred_logits": inp["det_logits"],
                "aux_outputs": [{"pred_boxes": det.detach(), "pred_logits": inp["det_logits"]}]}
    total = 0
    for box_type, sl in (("hand_boxes", slice(0, 2)), ("obj_boxes", slice(2, 4))):
        loss, idx = box_loss(detr_out, inp["det_px"][:, sl].clone(), box_type)
        total = total + loss
        res["box_" + box_type] = loss.detach()
        res["box_" + box_type + "_i"] = torch.cat([i for i, _ in idx])
        res["box_" + box_type + "_j"] = torch.cat([j for _, j in idx])
    total.backward()
    res["box_dpred"] = det.grad.clone()
    return res


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


def train_functional(case, forward, obj_proj, params):
    """Shared by the reference run (oracle/make_golden.py) and the oracle run: outputs of the training-mode forward and
    the gradients of a fixed random linear functional of (hs, all layers' boxes, obj_proj embeddings) w.r.t. every
    parameter (strided subsample + L2 norm per tensor).  forward(feats) -> (out, hs); params: name -> leaf tensor."""
    c = case["cfg"]
    feats, _ = make_inputs(case)
    g = torch.Generator().manual_seed(case["seed"] + 2000)
    out, hs = forward(feats)
    boxes = torch.stack([a["pred_boxes"] for a in out["aux_outputs"]] + [out["pred_boxes"]])
    emb = obj_proj(hs[-1])
    loss = (hs * torch.randn(hs.shape, generator=g)).sum() + (boxes * torch.randn(boxes.shape, generator=g)).sum() + \
        (emb * torch.randn(emb.shape, generator=g)).sum()
    loss.backward()
    res = {"hs": hs.detach().clone(), "boxes": boxes.detach().clone(), "embed": emb.detach().clone(),
           "loss": loss.detach().clone()}
    for k in sorted(params):
        gr = params[k].grad
        if gr is None:
            gr = torch.zeros_like(params[k])
        flat = gr.detach().flatten()
        res["grad/" + k] = flat[::max(1, flat.numel() // 64)][:64].clone()
        res["gnorm/" + k] = flat.norm().clone()
    return res


def run_oracle(case):
    """The restatement's answer for a case, in the same (subsampled) form as the fixture."""
    kind = case["kind"]
    if kind == "decoder_train":
        c = case["cfg"]
        sd = {k: v.clone().requires_grad_(True) for k, v in decoder_state_dict(case).items()}

        def fwd(feats):
            out, hs, _, _ = O.decoder_forward(feats, sd, heads=c["heads"], pred_traj=c["pred_traj"], dropout=case["dropout"])
            return out, hs
        return train_functional(case, fwd, lambda h: O.obj_proj(h, sd), sd)
    if kind == "losses":
        def word(nouns, pred, inds):
            loss, cols = O.word_contrastive_loss(nouns, pred, inds)
            return loss, torch.cat(cols)

        def matcher(outputs, targets, excl):
            tb, tl = [t["boxes"] for t in targets], [t["labels"] for t in targets]
            if excl:
                return O.hungarian_match(outputs["pred_boxes"], tb)
            return O.hungarian_match(outputs["pred_boxes"], tb, outputs["pred_logits"], tl)
        def box(detr_out, px, box_type):
            start, end = (0, 2) if box_type == "hand_boxes" else (2, 12)
            return O.box_loss(detr_out["pred_boxes"], px, start, end)
        return run_losses(make_inputs(case), O.sim_matrix,
                          lambda x, mv, mn, pad: O.egonce_loss(x, mv, mn, pad), word, matcher, box)
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


# ---- retrieval metrics: the reference's known-answer vector (utils/nDCG.py:154-172) and EPIC-MIR-like synthetic input
KNOWN_SIM = np.array([[1.0, 0.7, 0.4, 0.0], [0.3, 0.9, 0.6, 0.1], [0.2, 0.5, 0.8, 0.4]])
KNOWN_REL = np.array([[1.0, 0.5, 0.25, 0.0], [0.0, 1.0, 0.4, 0.0], [0.5, 0.3, 1.0, 0.0]])
KNOWN_K = np.array([[1, 1, 1, 0], [1, 1, 0, 0], [1, 1, 1, 0]])
KNOWN_NDCG = 0.9371789900735429            # utils/nDCG.py:172


def synth_retrieval(N, M, seed):
    """EPIC-MIR-like inputs: float64 similarities without ties, relevancies in {0, fractions, 1} with >= 1 exact one per row and column."""
    rng = np.random.RandomState(seed)
    sim = rng.randn(N, M)
    rel = np.where(rng.rand(N, M) < 0.02, np.round(rng.rand(N, M), 2), 0.0)
    rel[rng.rand(N, M) < 0.004] = 1.0
    rel[np.arange(N), rng.randint(0, M, N)] = 1.0
    rel[rng.randint(0, N, M), np.arange(M)] = 1.0          # ... in both retrieval directions
    return sim, rel

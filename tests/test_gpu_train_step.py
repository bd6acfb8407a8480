"""BASELINE config c4 in miniature: one training step of run/train.py:104-203 through the public mirrors -- frozen
backbone (video + 5 captions per clip) -> ObjDecoder -> txt_proj / obj_proj -> EgoNCE + hand / object box losses + word
loss -> backward -- against the same step written with the oracle's restatements under torch autograd (fp32 CPU)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import golden_cases as gc  # noqa: E402
from oracle import hh_oracle as O  # noqa: E402


def _inputs(B, T, img, R, V, g):
    video = torch.randn(B, T, 3, img, img, generator=g)
    tokens = gc.make_tokens(B * R, 128, g)
    pad = torch.ones(B * R)
    for i in range(B):                       # ~40 % of the rephrased captions are empty strings: [SOT, EOT, 0, ...]
        for r in range(1, R):
            if torch.rand(1, generator=g).item() < 0.4:
                tokens[i * R + r] = 0
                tokens[i * R + r, 0], tokens[i * R + r, 1] = 126, 127
                pad[i * R + r] = 0
    verb = (torch.rand(B, 12, generator=g) < 0.3).float()
    noun = (torch.rand(B, 20, generator=g) < 0.25).float()
    lo = 224 * torch.rand(B * T, 4, 2, generator=g) * 0.7
    px = torch.cat([lo, lo + 10 + 60 * torch.rand(B * T, 4, 2, generator=g)], -1)
    px[torch.rand(B * T, 4, generator=g) < 0.3] = 0.0
    noun_feats = torch.randn(V, 768, generator=g)
    inds = torch.randint(1, V, (B, 4), generator=g)
    inds[torch.rand(B, 4, generator=g) < 0.4] = 0
    inds[0, 0] = 3
    return dict(video=video, tokens=tokens, pad=pad[:, None].repeat(1, B), verb=verb, noun=noun, px=px,
                noun_feats=noun_feats, inds=inds)


@pytest.mark.parametrize("train_mode", [False, True])
def test_train_step_matches_oracle_autograd(train_mode):
    """train_mode: the decoder in train() as in run/train.py (dropout 0.1 at six sites per layer), the oracle graph with
    the same masks; otherwise eval() arithmetic."""
    drop = {"p": 0.1, "seed": 0xBEEF1234, "offset": 0} if train_mode else None
    from helping_hand_for_egocentric_videos_b200.model import LaviLa, box_utils, loss, metric, tfm_decoder as D
    B, T, img, R, V, Q = 4, 4, 56, 5, 40, 13
    enc = dict(img=img, patch=14, D=128, L=2, H=2, T=T, text_width=768, text_heads=12, text_layers=1, vocab=128)
    dec_cfg = dict(C=128, heads=2, layers=2, ffn=256, Q=Q, n=16, T=T, F=128, ncls=30, pred_traj=True)
    bsd = gc.backbone_state_dict(dict(cfg=enc, seed=101))
    dsd = gc.decoder_state_dict(dict(cfg=dec_cfg, seed=102))
    inp = _inputs(B, T, img, R, V, torch.Generator().manual_seed(103))
    eot = inp["tokens"].argmax(-1)
    wd = {"loss_bbox_hand_boxes": 5, "loss_bbox_obj_boxes": 5, "loss_giou_hand_boxes": 2, "loss_giou_obj_boxes": 2}

    # ---------------- oracle / autograd
    ref = {k: v.clone().requires_grad_(True) for k, v in dsd.items()}
    with torch.no_grad():
        bo = O.clip_forward(inp["video"], inp["tokens"], bsd, heads=2, text_heads=12)
    grid = bo["image_feature_map"][:, 1:].unflatten(1, (T, 16))
    out, hs, _, _ = O.decoder_forward(grid, ref, heads=2, pred_traj=True, dropout=drop)
    txt = O.txt_proj(bo["text_feature_map"][torch.arange(B * R), eot], ref)
    vid = O.obj_proj(hs[-1], ref)[:, -1]
    nce, _ = O.egonce_loss(O.sim_matrix(txt, vid), O.sim_matrix(inp["verb"], inp["verb"]),
                           O.sim_matrix(inp["noun"], inp["noun"]), inp["pad"])
    lh, _ = O.box_loss(out["pred_boxes"], inp["px"][:, :2], 0, 2)
    lo_, _ = O.box_loss(out["pred_boxes"], inp["px"][:, 2:], 2, 12)
    word, _ = O.word_contrastive_loss(O.txt_proj(inp["noun_feats"], ref), O.obj_proj(hs[-1], ref)[:, :-1], inp["inds"])
    total = nce + lh + lo_ + 0.5 * word
    total.backward()

    # ---------------- CUDA mirrors
    vis = LaviLa.SpaceTimeTransformer(img_size=img, patch_size=14, embed_dim=128, depth=2, num_heads=2, num_frames=T,
                                      time_init='zeros', ln_pre=True, act_layer=LaviLa.QuickGELU)
    vis.head = torch.nn.Identity()
    clip = LaviLa.CLIP(embed_dim=256, vision_width=128, vision_model=vis, context_length=77, vocab_size=128,
                       transformer_width=768, transformer_heads=12, transformer_layers=1)
    clip.load_state_dict(bsd, strict=True)
    clip = clip.cuda().eval()
    for p in clip.parameters():
        p.requires_grad = False                                                   # run/train.py:88 freezes the backbone
    tr = D.Cross_Attention(d_model=128, nhead=2, num_decoder_layers=2, dim_feedforward=256, normalize_before=True,
                           return_intermediate_dec=True)
    model = D.ObjDecoder(transformer=tr, num_classes=30, num_queries=Q, aux_loss=True, pred_traj=True, feature_dim=128,
                         num_frames=T, patches_per_frame=16)
    model.load_state_dict(dsd, strict=True)
    model = model.cuda().eval()
    if train_mode:
        model.train()
        model.dropout_seed, model._drop_step = drop["seed"], drop["offset"]
    crit = box_utils.SetCriterion(22047, matcher=box_utils.build_matcher(None), weight_dict=wd, eos_coef=0.1,
                                  losses=["boxes", "cardinality"]).cuda()
    dev = {k: v.cuda() for k, v in inp.items()}
    with torch.no_grad():
        o2 = clip(dev["video"], dev["tokens"], return_feature_map=True)
    grid2 = o2["image_feature_map"][:, 1:].unflatten(1, (T, 16))
    mo, hs2, _, _ = model(grid2)
    txt2 = model.txt_proj(o2["text_feature_map"][torch.arange(B * R), dev["tokens"].argmax(-1)])
    emb2 = model.obj_proj(hs2[-1])
    sim2 = metric.sim_matrix(txt2, emb2[:, -1].contiguous())
    nce2, _ = loss.EgoNCE()(sim2, metric.sim_matrix(dev["verb"], dev["verb"]), metric.sim_matrix(dev["noun"], dev["noun"]),
                            multi_pad_mask=dev["pad"], strict_mask=True)
    sizes = torch.full((B * T, 2), 224.0, device="cuda")
    lh2, _ = box_utils.compute_box_loss('hand_boxes', crit, mo, dev["px"][:, :2].clone(), None, sizes, n_queries=12)
    lo2, _ = box_utils.compute_box_loss('obj_boxes', crit, mo, dev["px"][:, 2:].clone(), None, sizes, n_queries=12)
    word2 = loss.WordContrastiveLoss()(model.txt_proj(dev["noun_feats"]), emb2[:, :-1].contiguous(), dev["inds"])
    total2 = nce2 + lh2 + lo2 + 0.5 * word2
    total2.backward()
    assert model.last_dropout == drop

    for a, b, nm in ((nce2, nce, "nce"), (lh2, lh, "hand"), (lo2, lo_, "obj"), (word2, word, "word")):
        assert abs(a.item() - b.item()) <= 2e-2 * max(1.0, abs(b.item())), (nm, a.item(), b.item())
    bad = []
    for k, p in model.named_parameters():
        want = ref[k].grad
        if want is None or want.abs().max() == 0:
            assert p.grad is None or p.grad.abs().max().item() <= 1e-6, k
            continue
        assert p.grad is not None, k
        got = p.grad.float().cpu().reshape(want.shape)
        cos = F.cosine_similarity(got.flatten(), want.flatten(), dim=0).item()
        if cos < 0.99:
            bad.append((k, round(cos, 4)))
    assert not bad, bad

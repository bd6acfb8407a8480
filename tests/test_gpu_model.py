"""Model-level parity (GPU): the drop-in modules (helping_hand_for_egocentric_videos_b200.model.*) against
  (a) the committed golden fixtures produced by the UNMODIFIED reference (tests/golden/*.pt), and
  (b) the oracle (oracle/hh_oracle.py, fp32 CPU) on identical seeded weights and inputs.

Gates are the ones BASELINE.json states for the bf16 path vs the fp32 reference: embedding cosine >= 0.999,
box L1 <= 1e-2, identical EgoMCQ argmax (checked where the fp32 margin exceeds the measured similarity error),
exact box / noun indices."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import golden_cases as gc  # noqa: E402
from oracle import hh_oracle as O  # noqa: E402


def _mods():
    from helping_hand_for_egocentric_videos_b200.model import LaviLa, tfm_decoder, metric
    return LaviLa, tfm_decoder, metric


def _cos(a, b):
    return F.cosine_similarity(a.float().flatten(1), b.float().flatten(1), dim=1).min().item()


def _build_backbone(case):
    LaviLa, _, _ = _mods()
    c = case["cfg"]
    vis = LaviLa.SpaceTimeTransformer(img_size=c["img"], patch_size=c["patch"], embed_dim=c["D"], depth=c["L"],
                                      num_heads=c["H"], num_frames=c["T"], time_init='zeros', ln_pre=True,
                                      act_layer=LaviLa.QuickGELU)
    vis.head = torch.nn.Identity()
    clip = LaviLa.CLIP(embed_dim=256, vision_width=c["D"], vision_model=vis, context_length=77, vocab_size=c["vocab"],
                       transformer_width=c["text_width"], transformer_heads=c["text_heads"],
                       transformer_layers=c["text_layers"])
    sd = gc.backbone_state_dict(case)
    clip.load_state_dict(sd, strict=True)          # key/shape contract of SURVEY.md section 8b
    return clip.cuda().eval(), sd


def _build_decoder(case):
    _, D, _ = _mods()
    c = case["cfg"]
    tr = D.Cross_Attention(d_model=c["C"], nhead=c["heads"], num_decoder_layers=c["layers"], dim_feedforward=c["ffn"],
                           normalize_before=True, return_intermediate_dec=True)
    dec = D.ObjDecoder(transformer=tr, num_classes=c["ncls"], num_queries=c["Q"], aux_loss=True, pred_traj=c["pred_traj"],
                       feature_dim=c["F"], num_frames=c["T"], patches_per_frame=c["n"])
    sd = gc.decoder_state_dict(case)
    dec.load_state_dict(sd, strict=True)
    return dec.cuda().eval(), sd


@pytest.mark.parametrize("name", ["enc_tiny", "enc_tiny_d1", "enc_c0", "enc_l14", "enc_l14_t16", "txt_large"])
def test_encoder_against_reference_golden(name):
    case = gc.CASES[name]
    ref = torch.load(os.path.join(gc.GOLDEN_DIR, name + ".pt"))
    clip, _ = _build_backbone(case)
    video, tokens = gc.make_inputs(case)
    out = clip(video.cuda(), tokens.cuda(), return_feature_map=True)
    got = gc.subsample(case, {k: out[k].cpu() for k in ref})
    assert got["image_feature_map"].shape == ref["image_feature_map"].shape
    assert _cos(got["image_embed"], ref["image_embed"]) >= 0.999
    assert _cos(got["image_feature_map"], ref["image_feature_map"]) >= 0.999
    err = (got["image_feature_map"] - ref["image_feature_map"]).abs().max().item()
    assert err <= (0.3 if name.startswith("enc_l14") else 0.15), err      # O(4) activations through 12 (L/14: 24) bf16 layers
    # text tower (hh_text_forward): bf16 GEMM operands, fp32 residual stream
    assert got["text_feature_map"].shape == ref["text_feature_map"].shape
    assert _cos(got["text_embed"], ref["text_embed"]) >= 0.999
    assert _cos(got["text_feature_map"], ref["text_feature_map"]) >= 0.999
    terr = (got["text_feature_map"] - ref["text_feature_map"]).abs().max().item()
    assert terr <= 0.08, terr


@pytest.mark.parametrize("width,heads,G", [(512, 8, 10), (768, 12, 3)])
def test_text_tower_full_depth_against_oracle(width, heads, G):
    """12-layer text tower at the BASE (512 x 8) and LARGE (768 x 12) geometry of reference model/LaviLa.py:55-170,
    vocabulary 49408, against the oracle's fp32 text_forward on the same seeded weights / captions."""
    LaviLa, _, _ = _mods()
    g = torch.Generator().manual_seed(width)
    shapes = {k: v for k, v in O.clip_param_shapes(128, 1, 16, 4, 1, text_width=width, text_layers=12,
                                                   vocab=49408).items() if not k.startswith("visual.")}
    sd = O.synth_state_dict(shapes, width)
    vis = LaviLa.SpaceTimeTransformer(img_size=32, patch_size=16, embed_dim=128, depth=1, num_heads=2, num_frames=1,
                                      time_init='zeros', ln_pre=True, act_layer=LaviLa.QuickGELU)
    vis.head = torch.nn.Identity()
    clip = LaviLa.CLIP(embed_dim=256, vision_width=128, vision_model=vis, context_length=77, vocab_size=49408,
                       transformer_width=width, transformer_heads=heads, transformer_layers=12)
    missing = clip.load_state_dict(sd, strict=False)
    assert all(k.startswith("visual.") for k in missing.missing_keys) and not missing.unexpected_keys
    clip = clip.cuda().eval()
    tokens = gc.make_tokens(G, 49408, g)
    x_cls, x = clip.encode_text(tokens.cuda())
    with torch.no_grad():
        want_cls, want_x = O.text_forward(tokens, sd, heads)
    assert _cos(x_cls.cpu(), want_cls) >= 0.999
    assert _cos(x.cpu(), want_x) >= 0.999
    err = (x.cpu() - want_x).abs().max().item()
    assert err <= 0.12, err
    # int32 ids are accepted; ids outside the vocabulary and CPU tensors fail loudly
    x_cls32, _ = clip.encode_text(tokens.int().cuda())
    assert torch.equal(x_cls32, x_cls)
    bad = tokens.clone()
    bad[0, 3] = 49408
    with pytest.raises(RuntimeError, match="token id"):
        clip.encode_text(bad.cuda())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        clip.encode_text(tokens)


@pytest.mark.parametrize("name", ["enc_tiny", "enc_c0"])
def test_uint8_frames_equal_host_normalised_clip(name):
    """the loader tail (base/base_dataset.py:322-323 `frames.float()/255`, `permute`;
    data_loader/transforms.py:48-51 NormalizeVideo with the constants of run/test_EgoMCQ.py:230-233) fused into the patch
    loader gives bit-identical features to normalising on the host and calling the fp32 entry point."""
    case = gc.CASES[name]
    clip, _ = _build_backbone(case)
    c = case["cfg"]
    g = torch.Generator().manual_seed(77)
    frames = torch.randint(0, 256, (2, c["T"], c["img"], c["img"], 3), generator=g, dtype=torch.uint8)
    mean = [108.3272985 / 255, 116.7460125 / 255, 104.09373615000001 / 255]
    std = [68.5005327 / 255, 66.6321579 / 255, 70.32316305 / 255]
    x = (frames.to(torch.float32) / 255).permute(0, 1, 4, 2, 3).contiguous()          # [B,T,3,H,W] on the host
    x = x.sub_(torch.tensor(mean)[None, None, :, None, None]).div_(torch.tensor(std)[None, None, :, None, None])
    cls_a, fmap_a = clip.visual.forward_features(x.cuda())
    cls_b, fmap_b = clip.visual.forward_features_u8(frames.cuda(), mean, std)
    assert torch.equal(fmap_a, fmap_b) and torch.equal(cls_a, cls_b)
    with pytest.raises(TypeError):
        clip.visual.forward_features_u8(frames.cuda().float(), mean, std)


def test_encoder_blocks_against_oracle():
    """Per-depth parity: truncated forwards (1, 2 blocks) against the oracle's per-block activations."""
    case = gc.CASES["enc_tiny"]
    clip, sd = _build_backbone(case)
    video, _ = gc.make_inputs(case)
    c = case["cfg"]
    with torch.no_grad():
        _, _, blocks = O.encoder_forward(video, sd, c["H"], pfx="visual.", return_blocks=True)
    for nb in (0, 1, 2):
        _, fmap = clip.visual.forward_features(video.cuda(), _nblocks=nb)
        x = O.encoder_embed(video, sd, "visual.") if nb == 0 else blocks[nb - 1]
        want = F.layer_norm(x, (c["D"],), sd["visual.norm.weight"], sd["visual.norm.bias"], 1e-6)
        err = (fmap.cpu() - want).abs().max().item()
        assert err <= (2e-2 if nb == 0 else 6e-2), (nb, err)     # nb=0: bf16 patch-embed GEMM only
        assert _cos(fmap.cpu(), want) >= 0.9995


@pytest.mark.parametrize("name", ["dec_tiny_traj", "dec_tiny_notraj", "dec_c0", "dec_c1", "dec_c2", "dec_c4"])
def test_decoder_against_reference_golden(name):
    case = gc.CASES[name]
    ref = torch.load(os.path.join(gc.GOLDEN_DIR, name + ".pt"))
    dec, _ = _build_decoder(case)
    _, _, metric = _mods()
    feats, text_feat = gc.make_inputs(case)
    out, hs, a, b = dec(feats.cuda())
    assert a == [] and b == []
    vid = dec.obj_proj(hs[-1])[:, -1]
    txt = dec.txt_proj(text_feat.cuda())
    res = {"pred_logits": out["pred_logits"], "pred_boxes": out["pred_boxes"], "hs": hs,
           "aux_boxes": torch.stack([x["pred_boxes"] for x in out["aux_outputs"]]),
           "aux_logits0": out["aux_outputs"][0]["pred_logits"], "video_embed": vid, "text_embed": txt,
           "sim": metric.sim_matrix(txt, vid)}
    got = gc.subsample(case, {k: v.cpu() for k, v in res.items()})
    for k in ref:
        assert got[k].shape == ref[k].shape, k
    assert (got["pred_boxes"] - ref["pred_boxes"]).abs().max() <= 1e-2            # box L1 gate
    assert (got["aux_boxes"] - ref["aux_boxes"]).abs().max() <= 1e-2
    assert _cos(got["video_embed"], ref["video_embed"]) >= 0.999                  # embedding gate
    assert _cos(got["hs"].flatten(0, 1), ref["hs"].flatten(0, 1)) >= 0.999
    assert torch.allclose(got["text_embed"], ref["text_embed"], atol=1e-4)        # fp32 head
    assert (got["sim"] - ref["sim"]).abs().max() <= 5e-3
    assert (got["pred_logits"] - ref["pred_logits"]).abs().max() <= 5e-2
    # class index parity where the reference's top-1 margin is above the logit error
    top2 = ref["pred_logits"].topk(2, -1).values
    sure = (top2[..., 0] - top2[..., 1]) > 0.1
    assert torch.equal(got["pred_logits"].argmax(-1)[sure], ref["pred_logits"].argmax(-1)[sure])


def test_decoder_accepts_the_strided_feature_map_view():
    """run/test_EgoMCQ.py:69-70 hands the decoder a view of image_feature_map[:, 1:] -- no copy needed."""
    case = gc.CASES["dec_tiny_traj"]
    dec, _ = _build_decoder(case)
    c = case["cfg"]
    feats, _ = gc.make_inputs(case)
    B = feats.shape[0]
    fmap = torch.zeros(B, 1 + c["T"] * c["n"], c["F"]).cuda()
    fmap[:, 1:] = feats.reshape(B, -1, c["F"]).cuda()
    view = fmap[:, 1:].unflatten(1, (c["T"], c["n"]))
    assert not view.is_contiguous()
    o1, hs1, _, _ = dec(view)
    o2, hs2, _, _ = dec(feats.cuda())
    assert torch.equal(hs1, hs2) and torch.equal(o1["pred_boxes"], o2["pred_boxes"])


def test_egomcq_end_to_end_choices():
    """EgoMCQ-style scoring (run/test_EgoMCQ.py:56-79) on BASELINE config c0 geometry, reduced depth for CPU time:
    G questions x 5 option clips + 1 caption; our choices vs the oracle's, margins logged."""
    LaviLa, D, metric = _mods()
    T, G = 4, 6
    enc_case = dict(kind="encoder", seed=77, B=5 * G, G=G,
                    cfg=dict(img=224, patch=16, D=768, L=2, H=12, T=T, text_width=768, text_heads=12, text_layers=1,
                             vocab=128))
    dec_case = dict(kind="decoder", seed=78, B=1, G=1,
                    cfg=dict(C=512, heads=8, layers=6, ffn=2048, Q=5, n=196, T=T, F=768, ncls=99, pred_traj=True))
    clip, bsd = _build_backbone(enc_case)
    dec, dsd = _build_decoder(dec_case)
    video, tokens = gc.make_inputs(enc_case)
    sims, ref_sims = [], []
    for g in range(G):
        v, t = video[5 * g:5 * g + 5], tokens[g:g + 1]
        out = clip(v.cuda(), t.cuda(), return_feature_map=True)
        grid = out['image_feature_map'][:, 1:].unflatten(1, (T, 196))
        txt = dec.txt_proj(out['text_feature_map'][0, t.argmax(-1).cuda()])
        _, hs, _, _ = dec(grid)
        vid = dec.obj_proj(hs[-1])[:, -1]
        sims.append(metric.sim_matrix(txt, vid).cpu())
        with torch.no_grad():
            ro = O.clip_forward(v, t, bsd, heads=12, text_heads=12)
            rgrid = ro['image_feature_map'][:, 1:].unflatten(1, (T, 196))
            _, rhs, _, _ = O.decoder_forward(rgrid, dsd, heads=8, pred_traj=True)
            ref_sims.append(O.sim_matrix(O.txt_proj(ro['text_feature_map'][0, t.argmax(-1)], dsd),
                                         O.obj_proj(rhs[-1], dsd)[:, -1]))
    sims, ref_sims = torch.stack(sims), torch.stack(ref_sims)          # [G,1,5]
    err = (sims - ref_sims).abs().max().item()
    top2 = ref_sims.reshape(G, 5).topk(2, -1).values
    margin = top2[:, 0] - top2[:, 1]
    mine = metric.egomcq_choices(sims.cuda()).cpu()
    ref = O.egomcq_choices(ref_sims)
    print("EgoMCQ sims max err %.2e ; margins %s ; choices %s vs %s" % (err, margin.tolist(), mine.tolist(), ref.tolist()))
    assert err <= 5e-3
    decided = margin > 2 * err
    assert torch.equal(mine[decided], ref[decided])
    labels = ref.clone()
    types = torch.tensor([1, 2] * (G // 2))
    acc = metric.egomcq_accuracy_metrics(sims, labels, types)
    if bool(decided.all()):
        assert acc == {"Intra-video": 100.0, "Inter-video": 100.0}


def test_state_dict_roundtrip_and_resync():
    """load_state_dict -> forward -> mutate a parameter in place -> forward sees the change (engine re-sync)."""
    case = gc.CASES["enc_tiny_d1"]
    clip, sd = _build_backbone(case)
    video, _ = gc.make_inputs(case)
    _, f1 = clip.visual.forward_features(video.cuda())
    _, f1b = clip.visual.forward_features(video.cuda())
    assert torch.equal(f1, f1b)
    with torch.no_grad():
        clip.visual.norm.weight.mul_(2.0)
    _, f2 = clip.visual.forward_features(video.cuda())
    w = sd["visual.norm.weight"].cuda()
    b = sd["visual.norm.bias"].cuda()
    assert torch.allclose((f1 - b) * 2 + b, f2, atol=1e-4)
    assert set(clip.state_dict().keys()) == set(sd.keys())


def test_cpu_input_fails_loudly():
    case = gc.CASES["enc_tiny_d1"]
    clip, _ = _build_backbone(case)
    video, _ = gc.make_inputs(case)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        clip.visual.forward_features(video)


@pytest.mark.parametrize("T,nq,traj", [(4, 4, True), (16, 12, False)])
def test_full_size_l14_against_oracle(T, nq, traj):
    """BASELINE configs c1 / c2 geometry at full depth and width (TimeSformer-L/14, 24 blocks, 1024 wide; decoder
    512 x 6 layers, 22 048 classes), one clip, against the fp32 oracle: the gates BASELINE.json states."""
    LaviLa, D, metric = _mods()
    from helping_hand_for_egocentric_videos_b200 import synthetic
    vis = LaviLa.SpaceTimeTransformer(img_size=224, patch_size=14, embed_dim=1024, depth=24, num_heads=16, num_frames=T,
                                      time_init='zeros', ln_pre=True, act_layer=LaviLa.QuickGELU, num_classes=0)
    tr = D.Cross_Attention(normalize_before=True, return_intermediate_dec=True)
    dec = D.ObjDecoder(tr, num_classes=22047, num_queries=nq + 1, aux_loss=True, pred_traj=traj, feature_dim=1024,
                       num_frames=T, patches_per_frame=256)
    synthetic.randomize_(vis, 5)
    synthetic.randomize_(dec, 6)
    vsd = {k: v.detach().clone() for k, v in vis.state_dict().items()}
    dsd = {k: v.detach().clone() for k, v in dec.state_dict().items()}
    vis, dec = vis.cuda().eval(), dec.cuda().eval()
    g = torch.Generator().manual_seed(8)
    video = torch.randn(1, T, 3, 224, 224, generator=g)
    text = torch.randn(5, 256, generator=g)
    _, fmap = vis.forward_features(video.cuda())
    out, hs, _, _ = dec(fmap[:, 1:].unflatten(1, (T, 256)))
    emb = dec.obj_proj(hs[-1])                      # [1, Q, 256]: hand / object / video query embeddings
    sim = metric.sim_matrix(text.cuda(), emb[0])
    with torch.no_grad():
        _, rfmap = O.encoder_forward(video, vsd, 16)
        rout, rhs, _, _ = O.decoder_forward(rfmap[:, 1:].unflatten(1, (T, 256)), dsd, heads=8, pred_traj=traj)
        remb = O.obj_proj(rhs[-1], dsd)
        rsim = O.sim_matrix(text, remb[0])
    cos_f = _cos(fmap.cpu(), rfmap)
    cos_e = F.cosine_similarity(emb.cpu()[0], remb[0], dim=-1).min().item()
    box = (out["pred_boxes"].cpu() - rout["pred_boxes"]).abs().max().item()
    serr = (sim.cpu() - rsim).abs().max().item()
    print("L/14 T=%d: fmap cos %.6f, embed cos %.6f, box L1 %.2e, sim err %.2e" % (T, cos_f, cos_e, box, serr))
    assert cos_f >= 0.999 and cos_e >= 0.999
    assert box <= 1e-2
    assert out["pred_logits"].shape == rout["pred_logits"].shape and out["pred_boxes"].shape == rout["pred_boxes"].shape
    assert serr <= 1e-2


def test_noun_and_box_indices_exact():
    """'box / noun index outputs exact' (SURVEY section 8a rows a17, a19): scipy's Hungarian assignment on costs built by
    our kernels equals the assignment on the oracle's costs -- noun matching uses -cos(noun, query) (model/loss.py:88-92)."""
    from scipy.optimize import linear_sum_assignment
    _, _, metric = _mods()
    g = torch.Generator().manual_seed(9)
    for trial in range(16):
        k = int(torch.randint(1, 5, (1,), generator=g))
        nouns = torch.randn(k, 256, generator=g)
        queries = torch.randn(12, 256, generator=g)
        mine = linear_sum_assignment((-metric.sim_matrix(nouns.cuda(), queries.cuda())).cpu().numpy())
        ref = linear_sum_assignment((-O.sim_matrix(nouns, queries)).numpy())
        assert (mine[1] == ref[1]).all()

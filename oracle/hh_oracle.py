"""CPU oracle: plain PyTorch fp32 restatement of the Helping-Hands video-side forward path.

TEST INFRASTRUCTURE ONLY -- never imported by the product package.  It exists so that the
CUDA path can be checked on the GPU box, where /root/reference is absent.

Parity status: PINNED.  `oracle/make_golden.py` runs the unmodified reference (imported through
`oracle/ref_import.py`) and this restatement on identical seeded state_dicts / inputs and commits
the reference outputs under tests/golden/; `tests/test_oracle_golden.py` re-checks the restatement
against those fixtures everywhere and against the live reference when /root/reference exists.

Everything here is a function of (state_dict, inputs); state_dict keys are exactly the reference's
(SURVEY.md §8b).  Citations are to files under /root/reference.
"""
from __future__ import annotations

import math
from typing import List, Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------------
# deterministic synthetic weights (shared by tests, smoke and bench; no reference code involved)
# --------------------------------------------------------------------------------------------

def encoder_param_shapes(D, L, patch, n, T, hidden=None):
    hidden = hidden or 4 * D
    s = {"cls_token": (1, 1, D), "pos_embed": (1, n + 1, D), "temporal_embed": (1, T, D),
         "patch_embed.proj.weight": (D, 3, patch, patch),
         "ln_pre.weight": (D,), "ln_pre.bias": (D,), "norm.weight": (D,), "norm.bias": (D,)}
    for i in range(L):
        p = "blocks.%d." % i
        for nm in ("norm1", "norm2", "norm3"):
            s[p + nm + ".weight"] = (D,)
            s[p + nm + ".bias"] = (D,)
        for at in ("attn", "timeattn"):
            s[p + at + ".qkv.weight"] = (3 * D, D)
            s[p + at + ".qkv.bias"] = (3 * D,)
            s[p + at + ".proj.weight"] = (D, D)
            s[p + at + ".proj.bias"] = (D,)
        s[p + "mlp.fc1.weight"] = (hidden, D)
        s[p + "mlp.fc1.bias"] = (hidden,)
        s[p + "mlp.fc2.weight"] = (D, hidden)
        s[p + "mlp.fc2.bias"] = (D,)
    return s


def decoder_param_shapes(C, Q, n, T, F_in, ncls1, layers=6, ffn=2048, pred_traj=True, text_width=768):
    s = {"pos_embed": (1, n + 1, C), "temporal_embed": (1, T, C),
         "txt_proj.1.weight": (256, text_width), "txt_proj.1.bias": (256,),
         "vid_proj.0.weight": (256, text_width), "vid_proj.0.bias": (256,),
         "transformer.pre_norm.weight": (C,), "transformer.pre_norm.bias": (C,),
         "transformer.decoder.norm.weight": (C,), "transformer.decoder.norm.bias": (C,),
         "class_embed.weight": (ncls1, C), "class_embed.bias": (ncls1,),
         "query_embed.weight": (Q, C), "proj.weight": (C, F_in),
         "obj_proj.0.weight": (C, C), "obj_proj.0.bias": (C,),
         "obj_proj.2.weight": (256, C), "obj_proj.2.bias": (256,)}
    dims = [C, C, C, 4]
    for j in range(3):
        s["bbox_embed.layers.%d.weight" % j] = (dims[j + 1], dims[j])
        s["bbox_embed.layers.%d.bias" % j] = (dims[j + 1],)
    if pred_traj:
        s["frame_index.weight"] = (T, C)
        s["frame_proj.weight"] = (C, 2 * C)
        s["frame_proj.bias"] = (C,)
    for i in range(layers):
        p = "transformer.decoder.layers.%d." % i
        for at in ("multihead_attn", "self_attn"):
            s[p + at + ".in_proj_weight"] = (3 * C, C)
            s[p + at + ".in_proj_bias"] = (3 * C,)
            s[p + at + ".out_proj.weight"] = (C, C)
            s[p + at + ".out_proj.bias"] = (C,)
        s[p + "linear1.weight"] = (ffn, C)
        s[p + "linear1.bias"] = (ffn,)
        s[p + "linear2.weight"] = (C, ffn)
        s[p + "linear2.bias"] = (C,)
        for nm in ("norm1", "norm2", "norm3"):
            s[p + nm + ".weight"] = (C,)
            s[p + nm + ".bias"] = (C,)
    return s


def clip_param_shapes(D, L, patch, n, T, text_width=768, text_layers=12, embed_dim=256, vocab=49408, ctx=77):
    """Full backbone state_dict: visual.* + text tower + projections (SURVEY.md §8b key list)."""
    s = {"visual." + k: v for k, v in encoder_param_shapes(D, L, patch, n, T).items()}
    W = text_width
    s.update({"positional_embedding": (ctx, W), "image_projection": (D, embed_dim), "text_projection": (W, embed_dim),
              "logit_scale": (), "token_embedding.weight": (vocab, W), "ln_final.weight": (W,), "ln_final.bias": (W,)})
    for i in range(text_layers):
        p = "transformer.resblocks.%d." % i
        s[p + "attn.in_proj_weight"] = (3 * W, W)
        s[p + "attn.in_proj_bias"] = (3 * W,)
        s[p + "attn.out_proj.weight"] = (W, W)
        s[p + "attn.out_proj.bias"] = (W,)
        for nm in ("ln_1", "ln_2"):
            s[p + nm + ".weight"] = (W,)
            s[p + nm + ".bias"] = (W,)
        s[p + "mlp.c_fc.weight"] = (4 * W, W)
        s[p + "mlp.c_fc.bias"] = (4 * W,)
        s[p + "mlp.c_proj.weight"] = (W, 4 * W)
        s[p + "mlp.c_proj.bias"] = (W,)
    return s


def synth_state_dict(shapes: Dict[str, tuple], seed: int) -> Dict[str, Tensor]:
    """Seeded synthetic weights.  Every tensor is re-randomised (the reference's `time_init='zeros'`
    would otherwise leave the temporal attention a no-op): matrices ~ N(0, fan_in^-1/2
    scaled), LayerNorm weights ~ 1 + 0.1 N(0,1), biases / embeddings ~ 0.02-0.1 N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k in sorted(shapes):
        shp = shapes[k]
        r = torch.randn(shp, generator=g, dtype=torch.float32)
        leaf = k.split(".")[-1]
        is_norm = ("norm" in k or "ln_" in k) and len(shp) == 1
        if is_norm and leaf == "weight":
            t = 1.0 + 0.1 * r
        elif is_norm:
            t = 0.05 * r
        elif len(shp) == 1:
            t = 0.05 * r
        elif k in ("query_embed.weight", "frame_index.weight"):
            t = 0.5 * r
        elif leaf in ("pos_embed", "temporal_embed", "cls_token", "positional_embedding") or k == "token_embedding.weight":
            t = 0.1 * r
        elif k == "logit_scale":
            t = torch.tensor(math.log(1 / 0.07))
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            t = r * (0.8 / math.sqrt(fan_in))
        sd[k] = t
    return sd


# --------------------------------------------------------------------------------------------
# encoder  (model/LaviLa.py)
# --------------------------------------------------------------------------------------------

def quick_gelu(x: Tensor) -> Tensor:
    """model/openai_model.py:177-179."""
    return x * torch.sigmoid(1.702 * x)


def _group_mask(T: int, n: int, mode: str) -> Tensor:
    """Boolean [N,N] attention mask equivalent to VarAttention's splice/rearrange/cat
    (model/LaviLa.py:255-276): query CLS sees every key; a patch query (f,p) sees the CLS key and
    the keys of its group -- same p over frames for 'time', same f over patches for 'space'."""
    N = 1 + T * n
    idx = torch.arange(T * n)
    f, p = idx // n, idx % n
    grp = p if mode == "time" else f
    m = torch.zeros(N, N, dtype=torch.bool)
    m[0, :] = True
    m[:, 0] = True
    m[1:, 1:] = grp[:, None] == grp[None, :]
    return m


def var_attention(z: Tensor, w_qkv, b_qkv, w_proj, b_proj, heads: int, T: int, n: int, mode: str) -> Tensor:
    """VarAttention.forward (model/LaviLa.py:246-283), grouped form: O(N (n + T)) like the reference, so that timing
    this oracle is a fair stand-in for the reference's CPU cost.  Checked against `var_attention_masked` in the tests."""
    B, N, D = z.shape
    hd = D // heads
    qkv = F.linear(z, w_qkv, b_qkv).view(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)      # [3,B,H,N,hd]
    q, k, v = qkv[0] * (hd ** -0.5), qkv[1], qkv[2]                                        # :249-252
    # CLS query: all 1 + T*n keys (:258)
    o_cls = torch.softmax(q[:, :, :1] @ k.transpose(-1, -2), dim=-1) @ v                   # [B,H,1,hd]
    # patch queries: {CLS key} U own group (:260-270)
    qp, kp, vp = (t[:, :, 1:].reshape(B, heads, T, n, hd) for t in (q, k, v))
    if mode == "time":          # group = same patch position over frames: put p before f
        qp, kp, vp = (t.transpose(2, 3) for t in (qp, kp, vp))                             # [B,H,n,T,hd]
    s_grp = qp @ kp.transpose(-1, -2)                                                      # [B,H,G,L,L]
    s_cls = (qp * k[:, :, :1, None]).sum(-1, keepdim=True)                                 # [B,H,G,L,1]
    p = torch.softmax(torch.cat([s_cls, s_grp], dim=-1), dim=-1)
    o = p[..., :1] * v[:, :, :1, None] + p[..., 1:] @ vp                                   # [B,H,G,L,hd]
    if mode == "time":
        o = o.transpose(2, 3)
    o = torch.cat([o_cls, o.reshape(B, heads, T * n, hd)], dim=2)                          # :276
    o = o.permute(0, 2, 1, 3).reshape(B, N, D)
    return F.linear(o, w_proj, b_proj)                                                     # :281


def var_attention_masked(z: Tensor, w_qkv, b_qkv, w_proj, b_proj, heads: int, T: int, n: int, mode: str) -> Tensor:
    """The same operator as masked dense attention (independent second statement, O(N^2))."""
    B, N, D = z.shape
    hd = D // heads
    qkv = F.linear(z, w_qkv, b_qkv).view(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * (hd ** -0.5), qkv[1], qkv[2]                      # :249-252
    s = q @ k.transpose(-1, -2)
    s = s.masked_fill(~_group_mask(T, n, mode).to(s.device), float("-inf"))
    o = torch.softmax(s, dim=-1) @ v                                       # :194-198
    o = o.permute(0, 2, 1, 3).reshape(B, N, D)
    return F.linear(o, w_proj, b_proj)                                     # :281


def encoder_block(x: Tensor, sd, pfx: str, heads: int, T: int, n: int, eps: float = 1e-6) -> Tensor:
    """SpaceTimeBlock.forward, 'frozen-in-time' wiring (model/LaviLa.py:345-390)."""
    D = x.shape[-1]
    g = lambda k: sd[pfx + k]
    ln = lambda t, nm: F.layer_norm(t, (D,), g(nm + ".weight"), g(nm + ".bias"), eps)
    t_out = var_attention(ln(x, "norm3"), g("timeattn.qkv.weight"), g("timeattn.qkv.bias"),
                          g("timeattn.proj.weight"), g("timeattn.proj.bias"), heads, T, n, "time")   # :353
    tr = x + t_out                                                                                    # :364
    s_out = var_attention(ln(tr, "norm1"), g("attn.qkv.weight"), g("attn.qkv.bias"),
                          g("attn.proj.weight"), g("attn.proj.bias"), heads, T, n, "space")          # :372
    sr = x + s_out                                                                                    # :384 (x, not tr)
    h = quick_gelu(F.linear(ln(sr, "norm2"), g("mlp.fc1.weight"), g("mlp.fc1.bias")))                 # :186-187
    return sr + F.linear(h, g("mlp.fc2.weight"), g("mlp.fc2.bias"))                                   # :388


def encoder_embed(video: Tensor, sd, pfx: str = "") -> Tensor:
    """Patch embed + CLS + pos/temporal embed + ln_pre (model/LaviLa.py:218-223,540-559)."""
    B, T, C, H, W = video.shape
    w = sd[pfx + "patch_embed.proj.weight"]
    D, _, p, _ = w.shape
    tok = F.conv2d(video.reshape(B * T, C, H, W), w, None, stride=p)         # no bias: ln_pre=True (:216)
    n = tok.shape[-1] * tok.shape[-2]
    tok = tok.flatten(2).transpose(1, 2).reshape(B, T * n, D)
    pos = sd[pfx + "pos_embed"]
    tpos = pos[:, 1:].repeat(1, T, 1) + sd[pfx + "temporal_embed"].repeat_interleave(n, 1)
    x = torch.cat([sd[pfx + "cls_token"].expand(B, -1, -1) + pos[:, :1], tok + tpos], 1)
    return F.layer_norm(x, (D,), sd[pfx + "ln_pre.weight"], sd[pfx + "ln_pre.bias"], 1e-5)            # :456 default eps


def encoder_forward(video: Tensor, sd, heads: int, pfx: str = "", depth: Optional[int] = None,
                    return_blocks: bool = False):
    """SpaceTimeTransformer.forward_features (model/LaviLa.py:537-573) -> (x_cls [B,D], fmap [B,N,D])."""
    B, T = video.shape[:2]
    x = encoder_embed(video, sd, pfx)
    D = x.shape[-1]
    n = (x.shape[1] - 1) // T
    if depth is None:
        depth = 1 + max(int(k[len(pfx):].split(".")[1]) for k in sd if k.startswith(pfx + "blocks."))
    blocks = []
    for i in range(depth):
        x = encoder_block(x, sd, "%sblocks.%d." % (pfx, i), heads, T, n)
        if return_blocks:
            blocks.append(x)
    fmap = F.layer_norm(x, (D,), sd[pfx + "norm.weight"], sd[pfx + "norm.bias"], 1e-6)               # :570,573
    if return_blocks:
        return fmap[:, 0], fmap, blocks
    return fmap[:, 0], fmap


def text_forward(tokens: Tensor, sd, heads: int) -> Tuple[Tensor, Tensor]:
    """CLIP.encode_text (model/LaviLa.py:660-670) + ResidualAttentionBlock (model/openai_model.py:182-216)."""
    x = sd["token_embedding.weight"][tokens] + sd["positional_embedding"]
    G, Lc, W = x.shape
    hd = W // heads
    causal = torch.full((Lc, Lc), float("-inf"), device=x.device).triu_(1)
    i = 0
    while ("transformer.resblocks.%d.ln_1.weight" % i) in sd:
        p = "transformer.resblocks.%d." % i
        y = F.layer_norm(x, (W,), sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], 1e-5)
        qkv = F.linear(y, sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"])
        qkv = qkv.view(G, Lc, 3, heads, hd).permute(2, 0, 3, 1, 4)
        s = (qkv[0] * hd ** -0.5) @ qkv[1].transpose(-1, -2) + causal
        o = (torch.softmax(s, -1) @ qkv[2]).permute(0, 2, 1, 3).reshape(G, Lc, W)
        x = x + F.linear(o, sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"])
        y = F.layer_norm(x, (W,), sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], 1e-5)
        x = x + F.linear(quick_gelu(F.linear(y, sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"])),
                         sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"])
        i += 1
    x = F.layer_norm(x, (W,), sd["ln_final.weight"], sd["ln_final.bias"], 1e-5)
    x_cls = x[torch.arange(G), tokens.argmax(-1)] @ sd["text_projection"]
    return x_cls, x


def clip_forward(video: Tensor, tokens: Optional[Tensor], sd, heads: int, text_heads: int = 12,
                 norm_embed: bool = True) -> Dict[str, Tensor]:
    """CLIP.forward(..., return_feature_map=True) (model/LaviLa.py:672-687)."""
    x_cls, fmap = encoder_forward(video, sd, heads, pfx="visual.")
    img = x_cls @ sd["image_projection"]                                                              # :657
    out = {"image_feature_map": fmap}
    if tokens is not None:
        txt, tmap = text_forward(tokens, sd, text_heads)
        out["text_feature_map"] = tmap
    else:
        txt = None
    if norm_embed:
        img = F.normalize(img, dim=-1)
        txt = F.normalize(txt, dim=-1) if txt is not None else None
    out["image_embed"] = img
    if txt is not None:
        out["text_embed"] = txt
    if "logit_scale" in sd:
        out["logit_scale"] = sd["logit_scale"].exp()
    return out


# --------------------------------------------------------------------------------------------
# object-aware decoder  (model/tfm_decoder.py)
# --------------------------------------------------------------------------------------------

def philox_keep(seed: int, offset: int, site: int, numel: int, p: float) -> Tensor:
    """Dropout multipliers (0 or 1/(1-p)) of one dropout site of the decoder's training forward, element order =
    row-major order of the dropped tensor.  Restates csrc/hh_rng.cuh (Philox4x32-10, key = seed, counter =
    (idx >> 3, site, offset), 16-bit lane idx & 7, keep <=> lane >= round(p * 65536)) so that the reference graph can be
    run with the masks the CUDA path draws.  The DISTRIBUTION is nn.Dropout's / F.dropout's (model/tfm_decoder.py:
    372-386; nn.MultiheadAttention(dropout=p) :365-366); the random stream is this project's, not torch's."""
    import numpy as np
    idx = np.arange(numel, dtype=np.uint64)
    blk = idx >> np.uint64(3)
    M32 = np.uint64(0xFFFFFFFF)
    c0 = blk & M32
    c1 = (blk >> np.uint64(32)) & M32
    c2 = np.full(numel, site, dtype=np.uint64)
    c3 = np.full(numel, offset, dtype=np.uint64)
    k0 = np.uint64(seed & 0xFFFFFFFF)
    k1 = np.uint64((seed >> 32) & 0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c0
        p1 = np.uint64(0xCD9E8D57) * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & M32
        hi1, lo1 = p1 >> np.uint64(32), p1 & M32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & M32, lo1, (hi0 ^ c3 ^ k1) & M32, lo0
        k0 = (k0 + np.uint64(0x9E3779B9)) & M32
        k1 = (k1 + np.uint64(0xBB67AE85)) & M32
    words = np.stack([c0, c1, c2, c3], 1)                                  # [numel, 4]
    lane = (idx & np.uint64(7)).astype(np.int64)
    w = words[np.arange(numel), lane >> 1]
    r = (w >> (np.uint64(16) * (lane & 1).astype(np.uint64))) & np.uint64(0xFFFF)
    thr = max(1, int(p * 65536.0 + 0.5))
    keep = (r >= np.uint64(thr)).astype(np.float32) * np.float32(1.0 / (1.0 - p))
    return torch.from_numpy(keep)


def _drop(x: Tensor, dropout, site: int) -> Tensor:
    """F.dropout(x, p, training=True) with the mask of philox_keep (dropout = None: eval mode, identity)."""
    if dropout is None or dropout["p"] <= 0:
        return x
    m = philox_keep(dropout["seed"], dropout["offset"], site, x.numel(), dropout["p"])
    return x * m.view(x.shape).to(x.dtype)


def _mha(q_in: Tensor, k_in: Tensor, v_in: Tensor, w, b, wo, bo, heads: int, dropout=None, site: int = 0) -> Tensor:
    """nn.MultiheadAttention forward (batch-first here), no masks (call sites model/tfm_decoder.py:433-441).
    dropout = None: eval mode; otherwise the attention probabilities are dropped after the softmax, as
    F.multi_head_attention_forward does in training mode."""
    C = q_in.shape[-1]
    hd = C // heads
    q = F.linear(q_in, w[:C], b[:C])
    k = F.linear(k_in, w[C:2 * C], b[C:2 * C])
    v = F.linear(v_in, w[2 * C:], b[2 * C:])
    B, Lq, _ = q.shape
    Lk = k.shape[1]
    q = q.view(B, Lq, heads, hd).transpose(1, 2) * (hd ** -0.5)
    k = k.view(B, Lk, heads, hd).transpose(1, 2)
    v = v.view(B, Lk, heads, hd).transpose(1, 2)
    o = _drop(torch.softmax(q @ k.transpose(-1, -2), -1), dropout, site) @ v           # probs [B, heads, Lq, Lk]
    return F.linear(o.transpose(1, 2).reshape(B, Lq, C), wo, bo)


def decoder_pos_embed(sd, T: int) -> Tensor:
    """ObjDecoder.construct_3d_pos_embed (model/tfm_decoder.py:161-166) -> [T*n, C]."""
    pe = sd["pos_embed"][0, 1:]
    te = sd["temporal_embed"][0]
    n = pe.shape[0]
    assert te.shape[0] == T, "reference requires T == num_frames (:161-166)"
    return (pe[None, :, :] + te[:, None, :]).reshape(T * n, -1)


def decoder_forward(features: Tensor, sd, heads: int = 8, pred_traj: bool = True,
                    num_frames: Optional[int] = None, dropout=None):
    """ObjDecoder.forward (model/tfm_decoder.py:183-233) with Cross_Attention.forward (:76-93),
    TransformerDecoder.forward (:255-295) and TransformerDecoderLayer.forward_pre (:420-461, sa_first).
    features [B,T,n,F] -> (out, hs[L,B,Q,C], [], []).
    dropout = {"p", "seed", "offset"}: training mode -- dropout1/2/3, the FFN's inner dropout (:372-386) and the
    attention-probability dropout of both nn.MultiheadAttention modules, with the masks of philox_keep
    (site = layer * 8 + {0 self-attn probs, 1 dropout1, 2 cross-attn probs, 3 dropout2, 4 FFN inner, 5 dropout3})."""
    B, T, n, _ = features.shape
    C = sd["proj.weight"].shape[0]
    mem = F.linear(features, sd["proj.weight"]).reshape(B, T * n, C)                                  # :200
    mem = F.layer_norm(mem, (C,), sd["transformer.pre_norm.weight"], sd["transformer.pre_norm.bias"], 1e-5)  # :86
    pos = decoder_pos_embed(sd, T)
    qpos = sd["query_embed.weight"][None].expand(B, -1, -1)
    tgt = torch.zeros_like(qpos)                                                                      # :84
    lnf = lambda t, p: F.layer_norm(t, (C,), sd[p + ".weight"], sd[p + ".bias"], 1e-5)
    hs = []
    L = 0
    while ("transformer.decoder.layers.%d.norm1.weight" % L) in sd:
        L += 1
    for i in range(L):
        p = "transformer.decoder.layers.%d." % i
        t2 = lnf(tgt, p + "norm1")
        tgt = tgt + _drop(_mha(t2 + qpos, t2 + qpos, t2, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"],
                               sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"], heads,
                               dropout, 8 * i + 0), dropout, 8 * i + 1)                              # :431-435
        t2 = lnf(tgt, p + "norm2")
        tgt = tgt + _drop(_mha(t2 + qpos, mem + pos, mem, sd[p + "multihead_attn.in_proj_weight"],
                               sd[p + "multihead_attn.in_proj_bias"], sd[p + "multihead_attn.out_proj.weight"],
                               sd[p + "multihead_attn.out_proj.bias"], heads, dropout, 8 * i + 2), dropout, 8 * i + 3)  # :438-456
        t2 = lnf(tgt, p + "norm3")
        ff = _drop(F.relu(F.linear(t2, sd[p + "linear1.weight"], sd[p + "linear1.bias"])), dropout, 8 * i + 4)
        tgt = tgt + _drop(F.linear(ff, sd[p + "linear2.weight"], sd[p + "linear2.bias"]), dropout, 8 * i + 5)   # :457-459
        hs.append(lnf(tgt, "transformer.decoder.norm"))                                               # :282
    hs = torch.stack(hs)                                                                              # [L,B,Q,C]
    logits = F.linear(hs, sd["class_embed.weight"], sd["class_embed.bias"])                           # :208
    nf = num_frames if num_frames is not None else sd["temporal_embed"].shape[1]
    if pred_traj and T == nf:
        Q = hs.shape[2]
        wf = sd["frame_proj.weight"]
        cond = F.linear(hs, wf[:, :C])[:, :, None] + \
            (F.linear(sd["frame_index.weight"], wf[:, C:], sd["frame_proj.bias"]))[None, None, :, None, :]
        cond = cond.reshape(L, B * T, Q, C)                                                           # :212-215
        logits = logits[:, :, None].expand(-1, -1, 4, -1, -1).flatten(1, 2)                           # :216 literal 4
    else:
        cond = hs
    x = cond
    for j in range(3):
        x = F.linear(x, sd["bbox_embed.layers.%d.weight" % j], sd["bbox_embed.layers.%d.bias" % j])
        if j < 2:
            x = F.relu(x)
    boxes = x.sigmoid()                                                                               # :228
    out = {"pred_logits": logits[-1], "pred_boxes": boxes[-1],
           "aux_outputs": [{"pred_logits": a, "pred_boxes": b} for a, b in zip(logits[:-1], boxes[:-1])]}
    return out, hs, [], []


def obj_proj(hs_last: Tensor, sd) -> Tensor:
    """ObjDecoder.obj_proj (model/tfm_decoder.py:175-180)."""
    return F.linear(F.relu(F.linear(hs_last, sd["obj_proj.0.weight"], sd["obj_proj.0.bias"])),
                    sd["obj_proj.2.weight"], sd["obj_proj.2.bias"])


def txt_proj(text_feat: Tensor, sd) -> Tensor:
    """ObjDecoder.txt_proj = ReLU -> Linear(768,256) (model/tfm_decoder.py:170-171)."""
    return F.linear(F.relu(text_feat), sd["txt_proj.1.weight"], sd["txt_proj.1.bias"])


# --------------------------------------------------------------------------------------------
# scoring and boxes  (model/metric.py, utils/box_ops.py, model/box_utils.py)
# --------------------------------------------------------------------------------------------

def sim_matrix(a: Tensor, b: Tensor, eps: float = 1e-8) -> Tensor:
    """model/metric.py:363-375."""
    a = a / a.norm(dim=-1, keepdim=True).clamp_min(eps)
    b = b / b.norm(dim=-1, keepdim=True).clamp_min(eps)
    return a @ b.transpose(-1, -2)


def egomcq_choices(sims: Tensor) -> Tensor:
    """argmax per question (model/metric.py:218): sims [G,1,5] or [G,5] -> int64 [G]."""
    return sims.reshape(sims.shape[0], -1).argmax(-1)


def egomcq_accuracy(preds: Tensor, labels: Tensor, types: Tensor) -> Dict[str, float]:
    """model/metric.py:209-225: accuracy per question type (sorted unique), first = Intra, second = Inter."""
    out = {}
    for t, name in zip(torch.unique(types).tolist(), ["Intra-video", "Inter-video"]):
        sel = types.reshape(-1) == t
        ch = egomcq_choices(preds[sel])
        out[name] = 100.0 * (ch == labels.reshape(-1)[sel]).float().mean().item()
    return out


def box_cxcywh_to_xyxy(x: Tensor) -> Tensor:
    """utils/box_ops.py:9-13."""
    cx, cy, w, h = x[..., 0], x[..., 1], x[..., 2], x[..., 3]
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], -1)


def box_xyxy_to_cxcywh(x: Tensor) -> Tensor:
    """utils/box_ops.py:16-20."""
    x0, y0, x1, y1 = x[..., 0], x[..., 1], x[..., 2], x[..., 3]
    return torch.stack([(x0 + x1) / 2, (y0 + y1) / 2, x1 - x0, y1 - y0], -1)


def box_iou(b1: Tensor, b2: Tensor) -> Tuple[Tensor, Tensor]:
    """utils/box_ops.py:24-37 -- note the +1e-4 on the union (:36)."""
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    lt = torch.maximum(b1[:, None, :2], b2[None, :, :2])
    rb = torch.minimum(b1[:, None, 2:], b2[None, :, 2:])
    wh = (rb - lt).clamp_min(0)
    inter = wh[..., 0] * wh[..., 1]
    union = a1[:, None] + a2[None, :] - inter
    return inter / (union + 0.0001), union


def generalized_box_iou(b1: Tensor, b2: Tensor) -> Tensor:
    """utils/box_ops.py:40-61 (xyxy in, [N,M] out)."""
    iou, union = box_iou(b1, b2)
    lt = torch.minimum(b1[:, None, :2], b2[None, :, :2])
    rb = torch.maximum(b1[:, None, 2:], b2[None, :, 2:])
    wh = (rb - lt).clamp_min(0)
    hull = wh[..., 0] * wh[..., 1]
    return iou - (hull - union) / hull


def matcher_cost(pred_cxcywh: Tensor, tgt_cxcywh: Tensor, w_bbox: float = 5.0, w_giou: float = 2.0) -> Tensor:
    """HungarianMatcher cost with exclude_class=True (model/box_utils.py:75-88, weights :96)."""
    l1 = torch.cdist(pred_cxcywh, tgt_cxcywh, p=1)
    g = generalized_box_iou(box_cxcywh_to_xyxy(pred_cxcywh), box_cxcywh_to_xyxy(tgt_cxcywh))
    return w_bbox * l1 + w_giou * (-g)


def egonce_logprobs(sim: Tensor, temperature: float = 0.07) -> Tuple[Tensor, Tensor]:
    """The softmax part of EgoNCE.forward (model/loss.py:61-69): log-softmax over rows and columns."""
    return F.log_softmax(sim / temperature, dim=1), F.log_softmax(sim.t() / temperature, dim=1)


# --------------------------------------------------------------------------------------------
# training-side matching / losses  (model/box_utils.py:43-92, model/loss.py:15-106)
# --------------------------------------------------------------------------------------------

def hungarian_match(pred_boxes: Tensor, tgt_boxes: List[Tensor], pred_logits: Optional[Tensor] = None,
                    tgt_labels: Optional[List[Tensor]] = None, cost_class: float = 1.0, cost_bbox: float = 5.0,
                    cost_giou: float = 2.0):
    """HungarianMatcher.forward (model/box_utils.py:43-92; weights of build_matcher :96).  `pred_logits`/`tgt_labels`
    given <=> exclude_class=False.  scipy.optimize.linear_sum_assignment is the reference's own solver (:6,91)."""
    from scipy.optimize import linear_sum_assignment
    bs, Q = pred_boxes.shape[:2]
    cost = matcher_cost(pred_boxes.flatten(0, 1), torch.cat(tgt_boxes), cost_bbox, cost_giou)
    if pred_logits is not None:
        prob = pred_logits.flatten(0, 1).softmax(-1)
        cost = cost + cost_class * -prob[:, torch.cat(tgt_labels)]
    cost = cost.view(bs, Q, -1)
    out, start = [], 0
    for i, t in enumerate(tgt_boxes):
        r, c = linear_sum_assignment(cost[i, :, start:start + len(t)])
        start += len(t)
        out.append((torch.as_tensor(r, dtype=torch.int64), torch.as_tensor(c, dtype=torch.int64)))
    return out


def egonce_loss(x: Tensor, mask_v: Optional[Tensor], mask_n: Optional[Tensor], multi_pad_mask: Optional[Tensor] = None,
                vn_threshold: float = 0.0, temperature: float = 0.07) -> Tuple[Tensor, Tensor]:
    """EgoNCE.forward (model/loss.py:15-70).  Caption row r belongs to video r // R (R = rows / columns); rows whose pad
    mask has a zero are filled with -inf and then dropped (:26,43-58); positives = (verb*noun | noun | verb) + diagonal,
    times the pad mask, thresholded (:59); loss = -mean_rows(mean_pos log_softmax_row) - mean_cols(...)  (:62-70)."""
    N, M = x.shape
    R = 1 if multi_pad_mask is None else N // M
    up = (lambda m: m) if R == 1 else (lambda m: m.repeat_interleave(R, 0))
    if mask_v is not None and mask_n is not None:
        extra = up(mask_v) * up(mask_n)
    else:
        extra = up(mask_n if mask_n is not None else mask_v)
    mask = extra + up(torch.eye(M, dtype=x.dtype))
    if multi_pad_mask is not None:
        mask = mask * multi_pad_mask
        keep = x.masked_fill(~multi_pad_mask.bool(), float("-inf")).sum(-1) != float("-inf")
        mask, x = mask[keep], x[keep]
    mb = mask > vn_threshold
    li = (F.log_softmax(x / temperature, dim=1) * mb).sum(1) / mb.sum(1)
    lj = (F.log_softmax(x.t() / temperature, dim=1) * mb.t()).sum(1) / mb.sum(0)
    return -li.mean() - lj.mean(), mb


def word_contrastive_loss(noun_embeds: Tensor, pred: Tensor, gt_inds: Tensor, temperature: float = 0.07,
                          noun_threshold: float = 0.6):
    """WordContrastiveLoss.forward (model/loss.py:78-106) -> (loss, [col_ind per clip with >= 1 noun])."""
    from scipy.optimize import linear_sum_assignment
    picked, cols = [], []
    for b in range(gt_inds.shape[0]):
        ids = gt_inds[b][gt_inds[b] != 0]
        if len(ids) == 0:
            continue
        cost = -sim_matrix(noun_embeds[ids], pred[b]).detach()          # :85-90
        _, col = linear_sum_assignment(cost)
        cols.append(torch.as_tensor(col, dtype=torch.int64))
        picked.append(pred[b][col])
    sel = torch.cat(picked)
    tgt = gt_inds[gt_inds != 0]
    logits = sim_matrix(sel, noun_embeds)                                # :96
    ns = sim_matrix(noun_embeds, noun_embeds).clone()
    ns.fill_diagonal_(0)                                                 # :99-100
    logits = logits.masked_fill(ns[tgt] > noun_threshold, -1) / temperature
    return F.cross_entropy(logits, tgt), cols


# --------------------------------------------------------------------------------------------
# retrieval metrics  (utils/mAP.py, utils/nDCG.py; numpy float64 like the reference)
# --------------------------------------------------------------------------------------------

def calculate_mAP(sim_mat, relevancy_matrix):
    """utils/mAP.py:4-44: AP_q = sum_k [rel_k == 1] * cumsum(rel)_k / (k+1) / #(rel == 1) over the ranking by descending
    similarity (note: the cumulative sum runs over ALL relevancy values, fractional ones included); mean over queries."""
    import numpy as np
    order = np.argsort(-sim_mat, axis=1, kind="stable")
    rows = np.arange(sim_mat.shape[0])[:, None]
    ranked = relevancy_matrix[rows, order]
    hits = ranked == 1
    cum = np.where(hits, np.cumsum(ranked, axis=1), 0)
    ap = np.sum(cum / (np.arange(ranked.shape[1]) + 1), axis=1) / hits.sum(axis=1)
    return np.mean(ap), ap


def calculate_DCG(similarity_matrix, relevancy_matrix, k_counts):
    """utils/nDCG.py:3-44: sum_k rel[rank_k] * k_counts[k] / log2(k + 2), ranking = argsort(sim)[::-1]."""
    import numpy as np
    order = np.argsort(similarity_matrix, axis=1, kind="stable")[:, ::-1]
    rows = np.arange(similarity_matrix.shape[0])[:, None]
    return np.sum(relevancy_matrix[rows, order] * k_counts / np.log2(np.arange(similarity_matrix.shape[1]) + 2), axis=1)


def calculate_k_counts(relevancy_matrix):
    """utils/nDCG.py:46-75."""
    import numpy as np
    return (np.sort(relevancy_matrix)[:, ::-1] > 0).astype(int)


def calculate_nDCG(similarity_matrix, relevancy_matrix, k_counts=None, IDCG=None):
    """utils/nDCG.py:97-150 with reduction='mean'; also returns the per-query vector."""
    import numpy as np
    if k_counts is None:
        k_counts = calculate_k_counts(relevancy_matrix)
    dcg = calculate_DCG(similarity_matrix, relevancy_matrix, k_counts)
    if IDCG is None:
        IDCG = calculate_DCG(relevancy_matrix, relevancy_matrix, k_counts)
    return np.mean(dcg / IDCG), dcg / IDCG


def prepare_targets(boxes: Tensor, center_crop_224: bool = True) -> List[Dict[str, Tensor]]:
    """prepare_targets(..., classes=None, center_crop=False) (model/box_utils.py:249-279): pixel xyxy boxes clipped to
    [0, 224] and normalised; all-zero / degenerate boxes dropped; labels = 1 - (box.sum != 0) (dummy, unused)."""
    labels = 1 - (boxes.sum(-1) != 0).float()
    b = torch.clip(boxes, min=0, max=224).div(224)
    out = []
    for c_, b_ in zip(labels, b):
        ok = (c_ != -1) * (b_[:, 2] > b_[:, 0]) * (b_[:, 3] > b_[:, 1])
        out.append({"labels": c_[ok], "boxes": box_xyxy_to_cxcywh(b_[ok])})
    return out


def box_loss(pred_boxes: Tensor, target_px: Tensor, start: int, end: int, w_bbox: float = 5.0, w_giou: float = 2.0):
    """compute_box_loss (model/box_utils.py:446-461) for one query range: split_detr_out [start, end), Hungarian matching
    with exclude_class=True, SetCriterion.loss_boxes (:157-173) = L1 / num_boxes and (1 - diag GIoU) / num_boxes, weighted
    5 / 2 and divided by len(weight_dict) / 3 = 4 / 3 (run/train.py:460-463)."""
    targets = prepare_targets(target_px)
    pred = pred_boxes[:, start:end]
    idx = hungarian_match(pred.detach(), [t["boxes"] for t in targets])
    num_boxes = max(float(sum(len(t["labels"]) for t in targets)), 1.0)
    src = torch.cat([pred[i, s] for i, (s, _) in enumerate(idx)])
    tgt = torch.cat([t["boxes"][j] for t, (_, j) in zip(targets, idx)])
    l1 = (src - tgt).abs().sum() / num_boxes
    giou = (1 - torch.diag(generalized_box_iou(box_cxcywh_to_xyxy(src), box_cxcywh_to_xyxy(tgt)))).sum() / num_boxes
    return (w_bbox * l1 + w_giou * giou) / (4 / 3), idx

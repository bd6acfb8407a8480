"""BASELINE config c4 on one GPU's share (64 clips x 4 frames, nq=12, 5 captions per clip): time of the frozen-backbone
forward, the decoder training forward, the three losses and the backward, with CUDA events.  Not a bench.py line (c4 is
a parity configuration); used to find the slow kernels of the training path.
  python tools/bench_train_step.py [B] [eval|train]      train (default): decoder in train() mode, dropout 0.1"""
import sys
import time

import torch

sys.path.insert(0, ".")
from helping_hand_for_egocentric_videos_b200 import synthetic  # noqa: E402
from helping_hand_for_egocentric_videos_b200.model import LaviLa, box_utils, loss, metric, tfm_decoder as D  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    T, R, V = 4, 5, 2000
    torch.cuda.set_device(0)
    clip = LaviLa.CLIP_OPENAI_TIMESFORMER_LARGE(num_frames=T)
    synthetic.randomize_(clip, 0)
    clip = clip.cuda().eval()
    for p in clip.parameters():
        p.requires_grad = False
    tr = D.Cross_Attention(normalize_before=True, return_intermediate_dec=True)
    model = D.ObjDecoder(transformer=tr, num_classes=22047, num_queries=13, aux_loss=True, pred_traj=True, feature_dim=1024,
                         num_frames=T, patches_per_frame=256)
    synthetic.randomize_(model, 1)
    model = model.cuda().eval()
    if len(sys.argv) <= 2 or sys.argv[2] != "eval":
        model.train()        # run/train.py trains the decoder in train() mode: dropout at six sites per layer
    print("decoder mode:", "train (dropout %.2f)" % tr.dropout_p if model.training else "eval", flush=True)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-5)
    crit = box_utils.SetCriterion(22047, matcher=box_utils.build_matcher(None), eos_coef=0.1, losses=["boxes", "cardinality"],
                                  weight_dict={"loss_bbox_hand_boxes": 5, "loss_bbox_obj_boxes": 5,
                                               "loss_giou_hand_boxes": 2, "loss_giou_obj_boxes": 2}).cuda()
    g = torch.Generator().manual_seed(5)
    video = torch.randn(B, T, 3, 224, 224, generator=g).cuda()
    tokens = torch.zeros(B * R, 77, dtype=torch.long)
    for i in range(B * R):
        ln = int(torch.randint(3, 30, (1,), generator=g))
        tokens[i, :ln] = torch.randint(1, 49405, (ln,), generator=g)
        tokens[i, ln] = 49407
    tokens = tokens.cuda()
    pad = (torch.rand(B * R, generator=g) > 0.4).float()
    pad[::R] = 1
    pad = pad[:, None].repeat(1, B).cuda()
    verb = (torch.rand(B, 118, generator=g) < 0.012).float().cuda()
    noun = (torch.rand(B, 582, generator=g) < 0.004).float().cuda()
    lo = 224 * torch.rand(B * T, 4, 2, generator=g) * 0.7
    px = torch.cat([lo, lo + 10 + 60 * torch.rand(B * T, 4, 2, generator=g)], -1)
    px[torch.rand(B * T, 4, generator=g) < 0.3] = 0.0
    px = px.cuda()
    noun_feats = torch.randn(V, 768, generator=g).cuda()
    inds = torch.randint(1, V, (B, 4), generator=g)
    inds[torch.rand(B, 4, generator=g) < 0.4] = 0
    inds = inds.cuda()
    sizes = torch.full((B * T, 2), 224.0, device="cuda")
    marks = {}

    def step(detail=False):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        host = {}

        def tick(name, t0):
            torch.cuda.synchronize()
            host[name] = (time.perf_counter() - t0) * 1e3
            return time.perf_counter()
        ev[0].record()
        with torch.no_grad():
            out = clip(video, tokens, return_feature_map=True)
        ev[1].record()
        grid = out["image_feature_map"][:, 1:].unflatten(1, (T, 256))
        mo, hs, _, _ = model(grid)
        ev[2].record()
        t0 = tick("_", time.perf_counter()) if detail else 0
        txt = model.txt_proj(out["text_feature_map"][torch.arange(B * R, device="cuda"), tokens.argmax(-1)])
        emb = model.obj_proj(hs[-1])
        nce, _ = loss.EgoNCE()(metric.sim_matrix(txt, emb[:, -1].contiguous()), metric.sim_matrix(verb, verb),
                               metric.sim_matrix(noun, noun), multi_pad_mask=pad, strict_mask=True)
        t0 = tick("  proj + EgoNCE", t0) if detail else 0
        lh, _ = box_utils.compute_box_loss('hand_boxes', crit, mo, px[:, :2].clone(), None, sizes, n_queries=12)
        t0 = tick("  box loss hands", t0) if detail else 0
        lo_, _ = box_utils.compute_box_loss('obj_boxes', crit, mo, px[:, 2:].clone(), None, sizes, n_queries=12)
        t0 = tick("  box loss objects", t0) if detail else 0
        word = loss.WordContrastiveLoss()(model.txt_proj(noun_feats), emb[:, :-1].contiguous(), inds)
        t0 = tick("  word loss", t0) if detail else 0
        total = nce + lh + lo_ + 0.5 * word
        ev[3].record()
        total.backward()
        ev[4].record()
        opt.step()
        opt.zero_grad(set_to_none=True)
        ev[5].record()
        torch.cuda.synchronize()
        names = ["backbone_fwd (video + %d captions)" % (B * R), "decoder_fwd_train", "heads + losses", "backward", "optimizer"]
        for i, nm in enumerate(names):
            marks[nm] = ev[i].elapsed_time(ev[i + 1])
        host.pop("_", None)
        marks.update(host)
        return float(total)

    for _ in range(2):
        step()
    t0 = time.perf_counter()
    n = 3
    for _ in range(n):
        lv = step()
    dt = (time.perf_counter() - t0) / n
    print("c4 share: %d clips x %d frames, loss %.4f, %.1f ms / step = %.1f clips/s" % (B, T, lv, dt * 1e3, B / dt))
    for k, v in marks.items():
        print("  %-40s %8.2f ms" % (k, v))
    step(detail=True)
    import cProfile
    import pstats
    pr = cProfile.Profile()
    out = clip(video[:1], tokens[:R], return_feature_map=True) if False else None
    with torch.no_grad():
        o = clip(video, tokens, return_feature_map=True)
    mo, hs, _, _ = model(o["image_feature_map"][:, 1:].unflatten(1, (T, 256)))
    torch.cuda.synchronize()
    pr.enable()
    box_utils.compute_box_loss('obj_boxes', crit, mo, px[:, 2:].clone(), None, sizes, n_queries=12)
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
    print("host-synchronised breakdown of 'heads + losses':")
    for k, v in marks.items():
        if k.startswith("  "):
            print("  %-40s %8.2f ms" % (k, v))


if __name__ == "__main__":
    main()

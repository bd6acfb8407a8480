"""Training-side rows of the path (a17 matcher, a18 EgoNCE, a19 word loss) on the GPU, through the
C ABI: against the reference's golden fixture (tests/golden/losses.pt, made by oracle/make_golden.py from the unmodified
reference incl. its autograd gradients), against the oracle at BASELINE c4 sizes, and -- for the assignment kernel --
against scipy (the reference's own solver) on random and heavily tied problems.

Tolerances: indices and masks exact; losses and gradients fp32 vs fp32 with different summation order: rel 2e-5."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import golden_cases as gc  # noqa: E402
from oracle import hh_oracle as O  # noqa: E402


def _mods():
    from helping_hand_for_egocentric_videos_b200 import ops
    from helping_hand_for_egocentric_videos_b200.model import box_utils, loss, metric
    return ops, box_utils, loss, metric


def _close(got, want, name, rtol=2e-5):
    got, want = got.detach().float().cpu(), want.detach().float().cpu()
    assert got.shape == want.shape, (name, got.shape, want.shape)
    err = (got - want).abs().max().item()
    assert err <= rtol * max(1.0, want.abs().max().item()), (name, err)


def _cuda_losses(inp):
    ops, box_utils, loss, metric = _mods()
    nce, wl, m = loss.EgoNCE(), loss.WordContrastiveLoss(), box_utils.build_matcher(None)
    dev = {k: ([t.cuda() for t in v] if isinstance(v, list) else v.cuda()) for k, v in inp.items()}

    def word(nouns, pred, inds):
        l = wl(nouns, pred, inds)
        cols = wl.last_col_ind.flatten()
        return l, cols[cols >= 0].cpu()

    def matcher(outputs, targets, excl):
        return m(outputs, targets, exclude_class=excl)
    crit = box_utils.SetCriterion(22047, matcher=m, eos_coef=0.1, losses=["boxes", "cardinality"],
                                  weight_dict={"loss_bbox_hand_boxes": 5, "loss_bbox_obj_boxes": 5,
                                               "loss_giou_hand_boxes": 2, "loss_giou_obj_boxes": 2}).cuda()

    def box(detr_out, px, box_type):
        sizes = torch.full((px.shape[0], 2), 224.0, device=px.device)
        return box_utils.compute_box_loss(box_type, crit, detr_out, px, None, sizes, n_queries=12)
    res = gc.run_losses(dev, metric.sim_matrix, lambda x, mv, mn, pad: nce(x, mv, mn, multi_pad_mask=pad, strict_mask=True),
                        word, matcher, box)
    return {k: v.cpu() for k, v in res.items()}


def test_losses_against_reference_golden():
    case = gc.CASES["losses"]
    ref = torch.load(os.path.join(gc.GOLDEN_DIR, "losses.pt"))
    got = _cuda_losses(gc.make_inputs(case))
    for k, v in ref.items():
        if v.dtype.is_floating_point:
            _close(got[k], v, k)
        else:
            assert torch.equal(got[k], v), k


@pytest.mark.parametrize("seed", [0, 1])
def test_assignment_matches_scipy(seed):
    """hh_assign == scipy.optimize.linear_sum_assignment, pair by pair, on wide / tall / square / empty problems, with
    row filters, and on small-integer costs where optimal assignments are massively non-unique (tie-break order)."""
    from scipy.optimize import linear_sum_assignment
    ops = _mods()[0]
    rng = np.random.RandomState(seed)
    P, R, Cc = 300, 13, 9
    cost = rng.randn(P, R, Cc).astype(np.float32)
    cost[P // 2:] = rng.randint(0, 4, size=(P - P // 2, R, Cc)).astype(np.float32)        # ties
    nr = rng.randint(0, R + 1, size=P)
    nc = rng.randint(0, Cc + 1, size=P)
    valid = rng.rand(P, R) < 0.7
    valid[::3] = True
    ri, ci, cnt = ops.assign(torch.from_numpy(cost).cuda(), [p * R * Cc for p in range(P)], [Cc] * P, nr, nc,
                             torch.from_numpy(valid).cuda())
    ri, ci, cnt = ri.cpu().numpy(), ci.cpu().numpy(), cnt.cpu().numpy()
    for p in range(P):
        rows = [r for r in range(nr[p]) if valid[p, r]]
        sub = cost[p][rows][:, :nc[p]].reshape(len(rows), nc[p])
        wr, wc = linear_sum_assignment(sub)
        k = cnt[p]
        assert k == len(wr), p
        assert (ri[p, :k] == wr).all() and (ci[p, :k] == wc).all(), (p, sub.shape)
        assert (ri[p, k:] == -1).all() and (ci[p, k:] == -1).all()


def test_assignment_reports_infeasible():
    ops = _mods()[0]
    cost = torch.full((2, 3, 3), float("inf")).cuda()
    cost[1] = torch.eye(3)
    _, _, cnt = ops.assign(cost, [0, 9], [3, 3], [3, 3], [3, 3])
    assert cnt.tolist() == [-1, 3]


@pytest.mark.parametrize("exclude_class", [True, False])
def test_matcher_at_c4_size(exclude_class):
    """HungarianMatcher on one GPU's share of BASELINE c4: 64 clips x 4 frames = 256 images, 10 object queries, 0-4 boxes."""
    _, box_utils, _, _ = _mods()
    g = torch.Generator().manual_seed(5)
    bs, Q, ncls = 256, 10, 22048

    def rb(k):
        return torch.cat([0.2 + 0.6 * torch.rand(k, 2, generator=g), 0.02 + 0.35 * torch.rand(k, 2, generator=g)], -1)
    pb = rb(bs * Q).view(bs, Q, 4)
    pl = torch.randn(bs, Q, ncls, generator=g)
    sizes = torch.randint(0, 5, (bs,), generator=g).tolist()
    tb = [rb(k) for k in sizes]
    tl = [torch.randint(0, ncls, (k,), generator=g) for k in sizes]
    m = box_utils.build_matcher(None)
    got = m({"pred_logits": pl.cuda(), "pred_boxes": pb.cuda()},
            [{"boxes": b.cuda(), "labels": l.cuda()} for b, l in zip(tb, tl)], exclude_class=exclude_class)
    want = O.hungarian_match(pb, tb) if exclude_class else O.hungarian_match(pb, tb, pl, tl)
    assert len(got) == bs
    for (gi, gj), (wi, wj) in zip(got, want):
        assert gi.dtype == torch.int64 and gi.device.type == "cpu"
        assert torch.equal(gi, wi) and torch.equal(gj, wj)


def test_matcher_without_targets_and_cpu_input():
    _, box_utils, _, _ = _mods()
    m = box_utils.build_matcher(None)
    out = {"pred_logits": torch.zeros(2, 5, 7).cuda(), "pred_boxes": torch.rand(2, 5, 4).cuda()}
    res = m(out, [{"boxes": torch.zeros(0, 4).cuda(), "labels": torch.zeros(0, dtype=torch.long).cuda()}] * 2, True)
    assert all(len(i) == 0 and len(j) == 0 for i, j in res)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m({k: v.cpu() for k, v in out.items()}, [{"boxes": torch.rand(1, 4), "labels": torch.zeros(1)}] * 2, True)


def test_egonce_at_c4_size():
    """[2560, 512] similarity matrix (8 GPUs x 64 clips, 5 captions each), ~40 % padded captions: loss, mask and the
    gradient that flows back into both embedding sets."""
    _, _, loss, metric = _mods()
    g = torch.Generator().manual_seed(9)
    Nv, R, d = 512, 5, 256
    vid = torch.randn(Nv, d, generator=g)
    txt = vid.repeat_interleave(R, 0) * 0.5 + torch.randn(Nv * R, d, generator=g)
    verb = (torch.rand(Nv, 118, generator=g) < 0.012).float()
    noun = (torch.rand(Nv, 582, generator=g) < 0.004).float()
    pad = (torch.rand(Nv * R, generator=g) > 0.4).float()
    pad[::R] = 1.0
    pad = pad[:, None].repeat(1, Nv)
    outs = []
    for dev in ("cpu", "cuda"):
        v, t = vid.clone().to(dev).requires_grad_(True), txt.clone().to(dev).requires_grad_(True)
        if dev == "cpu":
            sv, sn = O.sim_matrix(verb, verb), O.sim_matrix(noun, noun)
            l, mb = O.egonce_loss(O.sim_matrix(t, v), sv, sn, pad)
        else:
            sv, sn = metric.sim_matrix(verb.cuda(), verb.cuda()), metric.sim_matrix(noun.cuda(), noun.cuda())
            l, mb = loss.EgoNCE()(metric.sim_matrix(t, v), sv, sn, multi_pad_mask=pad.cuda(), strict_mask=True)
        l.backward()
        outs.append((l.detach().cpu(), mb.cpu(), v.grad.cpu(), t.grad.cpu()))
    (l0, m0, gv0, gt0), (l1, m1, gv1, gt1) = outs
    assert torch.equal(m0, m1)
    _close(l1, l0, "loss")
    _close(gv1, gv0, "d video", rtol=1e-4)
    _close(gt1, gt0, "d text", rtol=1e-4)


def test_word_loss_at_c4_size():
    _, _, loss, _ = _mods()
    g = torch.Generator().manual_seed(10)
    V, d, B2, Q, Wm = 2000, 256, 64, 12, 4
    nouns = torch.randn(V, d, generator=g)
    nouns[1::50] = nouns[0::50] + 0.2 * torch.randn(V // 50, d, generator=g)      # synonyms
    pred = torch.randn(B2, Q, d, generator=g)
    inds = torch.randint(1, V, (B2, Wm), generator=g)
    inds[torch.rand(B2, Wm, generator=g) < 0.4] = 0
    inds[0] = torch.tensor([50, 51, 0, 100])
    n0, p0 = nouns.clone().requires_grad_(True), pred.clone().requires_grad_(True)
    l0, cols0 = O.word_contrastive_loss(n0, p0, inds)
    l0.backward()
    n1, p1 = nouns.cuda().requires_grad_(True), pred.cuda().requires_grad_(True)
    wl = loss.WordContrastiveLoss()
    l1 = wl(n1, p1, inds.cuda())
    l1.backward()
    cols1 = wl.last_col_ind.cpu()
    assert torch.equal(cols1[cols1 >= 0], torch.cat(cols0))               # noun indices exact
    assert torch.equal(cols1 >= 0, inds != 0)
    _close(l1, l0, "loss")
    _close(n1.grad, n0.grad, "d nouns", rtol=1e-4)
    _close(p1.grad, p0.grad, "d pred", rtol=1e-4)


def test_sim_matrix_backward_with_clamped_rows():
    ops = _mods()[0]
    g = torch.Generator().manual_seed(2)
    a = torch.randn(37, 48, generator=g)
    b = torch.randn(21, 48, generator=g)
    a[3] = 0.0                                     # norm below eps: the clamp blocks the norm's gradient
    G = torch.randn(37, 21, generator=g)
    a0, b0 = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    O.sim_matrix(a0, b0).backward(G)
    da, db = ops.sim_matrix_backward(a.cuda(), b.cuda(), G.cuda())
    _close(da[torch.arange(37) != 3], a0.grad[torch.arange(37) != 3], "da")
    _close(db, b0.grad, "db", rtol=1e-4)
    assert torch.isfinite(da).all()


# ------------------------------------------------------------------------------------------ retrieval metrics (f-4)
def test_retrieval_known_answer_and_oracle():
    """utils/nDCG.py:154-181 known answer, then mAP / nDCG against the oracle at growing sizes up to one EPIC-MIR-sized
    row length (9668 candidates).  float64, numpy's summation order: equality is exact."""
    from helping_hand_for_egocentric_videos_b200.utils import mAP, nDCG
    from oracle.golden_cases import KNOWN_K, KNOWN_NDCG, KNOWN_REL, KNOWN_SIM, synth_retrieval
    assert (nDCG.calculate_k_counts(KNOWN_REL) == KNOWN_K).all()
    assert nDCG.calculate_nDCG(KNOWN_SIM, KNOWN_REL, KNOWN_K) == KNOWN_NDCG
    idcg = nDCG.calculate_IDCG(KNOWN_REL, KNOWN_K)
    assert nDCG.calculate_nDCG(KNOWN_SIM, KNOWN_REL, KNOWN_K, IDCG=idcg) == KNOWN_NDCG
    assert np.mean(nDCG.calculate_nDCG(KNOWN_SIM, KNOWN_REL, KNOWN_K, IDCG=idcg, reduction=None)) == KNOWN_NDCG
    for N, M, seed in [(5, 7, 0), (40, 300, 1), (33, 2000, 2), (24, 9668, 3)]:
        sim, rel = synth_retrieval(N, M, seed)
        want_map, want_ap = O.calculate_mAP(sim, rel)
        want_ndcg, want_rows = O.calculate_nDCG(sim, rel)
        assert mAP.calculate_mAP(sim, rel) == want_map, (N, M)
        assert np.array_equal(nDCG.calculate_nDCG(sim, rel, reduction=None), want_rows), (N, M)
        assert nDCG.calculate_nDCG(sim, rel) == want_ndcg
        # the text->video direction of run/test_epic.py:271-278 (transposed, non-contiguous inputs)
        assert mAP.calculate_mAP(sim.T, rel.T) == O.calculate_mAP(sim.T.copy(), rel.T.copy())[0]

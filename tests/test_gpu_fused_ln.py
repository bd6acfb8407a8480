"""LayerNorm folded into the tcgen05 GEMM epilogues (reference: norm1 / norm2 / norm3 and the residual adds of
SpaceTimeBlock.forward, model/LaviLa.py:353-388), through the C ABI, against plain fp32 PyTorch on the same bf16 operands.

  producer  hh_gemm_bf16_res_stats   z = A W^T + b + residual (bf16), per-row (sum, sum^2) partials, optional fp32 write-back
  consumer  hh_gemm_bf16_ln          act(Linear(LayerNorm(z))) = act(rstd (z W'^T - mean colsum) + b')
  fold      hh_fold_layernorm_weight W' = W diag(gamma), colsum, b' = b + W beta

Tolerances: bf16 output rounding (rel 2^-7) on top of fp32 accumulation; the fp32 write-back and the statistics are held
to 1e-4 relative."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import hh_oracle as O  # noqa: E402


def _ops():
    from helping_hand_for_egocentric_videos_b200 import ops
    return ops


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


def _close(got, want, rtol, atol, name=""):
    got, want = got.float(), want.float()
    err = (got - want).abs()
    bad = err > atol + rtol * want.abs()
    assert not bad.any(), "%s: %d/%d out of tolerance, max err %.3e (ref scale %.3e)" % (
        name, int(bad.sum()), bad.numel(), err.max().item(), want.abs().max().item())


def test_fold_layernorm_weight():
    ops = _ops()
    N, K = 3 * 256, 256
    w = _rand(N, K, seed=1, scale=1 / 16)
    gamma = 1 + 0.2 * _rand(K, seed=2)
    beta = 0.3 * _rand(K, seed=3)
    bias = _rand(N, seed=4, scale=0.5)
    wf, cs, bf = ops.fold_layernorm_weight(w, gamma, beta, bias, scaled_rows=K, scale=0.125)
    sc = torch.ones(N, 1, device="cuda")
    sc[:K] = 0.125
    want_w = (w * gamma[None, :] * sc).bfloat16()
    assert torch.equal(wf, want_w)
    _close(cs, want_w.float().sum(1), 1e-5, 1e-5, "colsum")
    _close(bf, (bias + w @ beta) * sc[:, 0], 1e-5, 1e-5, "folded bias")
    wf2, cs2, bf2 = ops.fold_layernorm_weight(w, gamma, beta, None)
    _close(bf2, w @ beta, 1e-5, 1e-5, "folded bias (no bias)")


# (M, N, K): small M -> 128-wide unclustered tiles; M >= 74 row tiles -> CTA pairs on 256-wide 2-SM tiles; ragged tails;
# K = 4096 is the fc2 shape, N = 768 / 256 the B/16 and smoke widths
RES_SHAPES = [(128, 128, 64), (300, 1024, 1024), (4097, 1024, 1024), (515, 1024, 4096), (785, 768, 768), (97, 256, 1024),
              (9473, 1024, 1024), (9601, 1024, 256), (12000, 768, 3072), (20000, 1024, 128), (9500, 256, 64),
              (9472, 640, 128)]


@pytest.mark.parametrize("M,N,K", RES_SHAPES)
@pytest.mark.parametrize("writeback", [False, True])
def test_gemm_res_stats(M, N, K, writeback):
    ops = _ops()
    a = _rand(M, K, seed=1).bfloat16()
    w = _rand(N, K, seed=2, scale=1 / math.sqrt(K)).bfloat16()
    bias = _rand(N, seed=3, scale=0.5)
    x = _rand(M, N, seed=4, scale=2.0) + 0.3
    x0 = x.clone()
    want = (a.float() @ w.float().t() + bias) + x0
    z, stats = ops.gemm_bf16_res_stats(a, w, bias, x, writeback=writeback)
    _close(z, want, 2 ** -7, 2e-3, "z")
    if writeback:
        _close(x, want, 1e-4, 2e-4, "fp32 write-back")
    else:
        assert torch.equal(x, x0), "residual must stay untouched without write-back"
    s = stats.sum(0)
    _close(s[:, 0], want.sum(1), 1e-4, 1e-3 * math.sqrt(N), "row sums")
    _close(s[:, 1], (want * want).sum(1), 1e-4, 1e-3, "row sums of squares")


def test_gemm_res_stats_no_bias_and_determinism():
    ops = _ops()
    M, N, K = 9600, 1024, 1024
    a = _rand(M, K, seed=5).bfloat16()
    w = _rand(N, K, seed=6, scale=1 / 32).bfloat16()
    x = _rand(M, N, seed=7)
    z1, s1 = ops.gemm_bf16_res_stats(a, w, None, x.clone(), writeback=True)
    z2, s2 = ops.gemm_bf16_res_stats(a, w, None, x.clone(), writeback=True)
    assert torch.equal(z1, z2) and torch.equal(s1, s2)
    _close(z1, a.float() @ w.float().t() + x, 2 ** -7, 2e-3, "z (no bias)")


LN_SHAPES = [(128, 256, 128), (300, 3072, 1024), (4097, 3072, 1024), (1000, 4096, 1024), (785, 2304, 768),
             (97, 768, 256), (9473, 3072, 1024), (9601, 4096, 1024), (12000, 512, 256), (20000, 384, 128)]


@pytest.mark.parametrize("M,N,K", LN_SHAPES)
@pytest.mark.parametrize("qgelu", [False, True])
@pytest.mark.parametrize("parts", [1, 4])
def test_gemm_ln(M, N, K, qgelu, parts):
    ops = _ops()
    zf = _rand(M, K, seed=1, scale=1.5) + 0.4 * _rand(M, 1, seed=8)          # rows with different, non-zero means
    z = zf.bfloat16()
    w = _rand(N, K, seed=2, scale=1 / math.sqrt(K))
    gamma = 1 + 0.2 * _rand(K, seed=3)
    beta = 0.2 * _rand(K, seed=4)
    bias = _rand(N, seed=5, scale=0.5)
    wf, cs, bf = ops.fold_layernorm_weight(w, gamma, beta, bias)
    # statistics of the rows the GEMM multiplies (the bf16 copy), split into `parts` column blocks like the producer's
    zz = z.float()
    blocks = zz.chunk(parts, dim=1)
    stats = torch.stack([torch.stack([b.sum(1), (b * b).sum(1)], -1) for b in blocks], 0).contiguous()
    for eps in (1e-6, 1e-5):
        got = ops.gemm_bf16_ln(z, wf, bf, cs, stats, eps, qgelu=qgelu)
        ref = F.layer_norm(zz, (K,), gamma, beta, eps) @ w.t() + bias
        if qgelu:
            ref = O.quick_gelu(ref)
        # the folded form rounds W gamma (not W) and z (not LN(z)) to bf16: same 2^-9 relative operand error
        _close(got, ref, 2 ** -6, 2e-2, "folded LayerNorm GEMM")
        err = (got.float() - ref).abs().mean().item() / ref.abs().mean().item()
        assert err < 4e-3, err


def test_producer_consumer_chain_matches_unfused_kernels():
    """proj -> (+x) -> LayerNorm -> fc1 through the fused pair against the stand-alone kernels of the same library."""
    ops = _ops()
    M, D, Hd = 9700, 1024, 4096
    a = _rand(M, D, seed=1).bfloat16()
    wp = _rand(D, D, seed=2, scale=1 / 32).bfloat16()
    bp = _rand(D, seed=3, scale=0.1)
    x = _rand(M, D, seed=4)
    w1 = _rand(Hd, D, seed=5, scale=1 / 32)
    b1 = _rand(Hd, seed=6, scale=0.1)
    gamma = 1 + 0.1 * _rand(D, seed=7)
    beta = 0.1 * _rand(D, seed=8)
    # fused
    xf = x.clone()
    z, stats = ops.gemm_bf16_res_stats(a, wp, bp, xf, writeback=True)
    wf, cs, bf = ops.fold_layernorm_weight(w1, gamma, beta, b1)
    h = ops.gemm_bf16_ln(z, wf, bf, cs, stats, 1e-6, qgelu=True)
    # fp32 statement
    xs = x + (a.float() @ wp.float().t() + bp)
    ref = O.quick_gelu(F.layer_norm(xs, (D,), gamma, beta, 1e-6) @ w1.t() + b1)
    _close(xf, xs, 1e-4, 2e-4, "x")
    rel = (h.float() - ref).abs().mean().item() / ref.abs().mean().item()
    assert rel < 5e-3, rel
    # un-fused kernels (round-1 sequence)
    dl = ops.gemm_bf16(a, wp, bp, epilogue=0)
    _, a16 = ops.layernorm(x + dl.float(), gamma, beta, 1e-6, want_f32=False, want_bf16=True)
    h0 = ops.gemm_bf16(a16, w1.bfloat16(), b1, epilogue=1)
    rel0 = (h0.float() - ref).abs().mean().item() / ref.abs().mean().item()
    assert rel < 1.5 * rel0 + 1e-4, (rel, rel0)


def _encoder(fused, T=4, depth=3, D=256, H=4, img=56):
    from helping_hand_for_egocentric_videos_b200.model import LaviLa
    from helping_hand_for_egocentric_videos_b200 import synthetic
    old = os.environ.get("HH_LN_UNFUSED")
    os.environ["HH_LN_UNFUSED"] = "0" if fused else "1"
    try:
        vis = LaviLa.SpaceTimeTransformer(img_size=img, patch_size=14, embed_dim=D, depth=depth, num_heads=H, num_frames=T,
                                          time_init='zeros', ln_pre=True, act_layer=LaviLa.QuickGELU, num_classes=0)
        synthetic.randomize_(vis, 11)
        vis = vis.cuda().eval()
        vis._engine()            # the engine reads the switch at construction
    finally:
        if old is None:
            del os.environ["HH_LN_UNFUSED"]
        else:
            os.environ["HH_LN_UNFUSED"] = old
    return vis


@pytest.mark.parametrize("B", [1, 3, 40])
def test_encoder_fused_vs_unfused_and_oracle(B):
    """Whole encoder: folded-LayerNorm path vs the round-1 kernel sequence vs the fp32 oracle."""
    from helping_hand_for_egocentric_videos_b200 import synthetic
    T = 4
    fused, plain = _encoder(True, T), _encoder(False, T)
    video = synthetic.synthetic_clips(B, T, 56, seed=5, device="cuda")
    _, f1 = fused.forward_features(video)
    _, f0 = plain.forward_features(video)
    sd = {k: v.detach().cpu() for k, v in fused.state_dict().items()}
    with torch.no_grad():
        _, ref = O.encoder_forward(video.cpu(), sd, 4)
    ref = ref.cuda()
    cos1 = F.cosine_similarity(f1.flatten(1), ref.flatten(1), dim=1).min().item()
    cos0 = F.cosine_similarity(f0.flatten(1), ref.flatten(1), dim=1).min().item()
    e1 = (f1 - ref).abs().mean().item()
    e0 = (f0 - ref).abs().mean().item()
    assert cos1 >= 0.9995 and cos0 >= 0.9995, (cos1, cos0)
    assert e1 <= 1.5 * e0 + 1e-4, (e1, e0)          # folding must not cost accuracy
    assert fused.last_launches() < plain.last_launches()

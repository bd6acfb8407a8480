"""Backward kernels of the decoder against torch autograd on the same fp32 statement of
each operator, through the C ABI.  fp32 kernels: rel 2e-5 (different summation order); the cross-attention backward
reads bf16 K/V and writes bf16 dK/dV: compared on the same bf16-rounded inputs, output rounding 2^-8."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ops():
    from helping_hand_for_egocentric_videos_b200 import ops
    return ops


def _close(got, want, name, rtol=2e-5, atol=None):
    got, want = got.detach().float().cpu(), want.detach().float().cpu()
    assert got.shape == want.shape, (name, got.shape, want.shape)
    err = (got - want).abs().max().item()
    lim = (atol if atol is not None else 0.0) + rtol * max(1.0, want.abs().max().item())
    assert err <= lim, (name, err, lim)


@pytest.mark.parametrize("R,N,K,act,add,in_relu", [(65, 512, 512, 0, True, False), (832, 2048, 512, 1, False, False),
                                                   (100, 4, 512, 2, False, False), (37, 256, 768, 0, False, True),
                                                   (5, 40, 24, 1, True, False)])
def test_linear_backward(R, N, K, act, add, in_relu):
    g = torch.Generator().manual_seed(R + N)
    x = torch.randn(R, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g) * 0.1
    xa = torch.randn(13 if R >= 13 else R, K, generator=g) if add else None
    dy = torch.randn(R, N, generator=g)
    x0, w0, b0 = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    xin = x0 + (xa.repeat((R + xa.shape[0] - 1) // xa.shape[0], 1)[:R] if add else 0)
    if in_relu:
        xin = F.relu(xin)
    y = F.linear(xin, w0, b0)
    y = F.relu(y) if act == 1 else (torch.sigmoid(y) if act == 2 else y)
    y.backward(dy)
    dx, dw, db = _ops().linear_f32_backward(dy.cuda(), y.detach().cuda(), act, w.cuda(), x.cuda(),
                                            xa.cuda() if add else None, in_relu)
    _close(dw, w0.grad, "dw", 5e-5)
    _close(db, b0.grad, "db", 5e-5)
    if not in_relu:          # the kernel returns d(relu(x)) there; the caller applies the mask
        _close(dx, x0.grad, "dx", 5e-5)


@pytest.mark.parametrize("M,D", [(832, 512), (3000, 1024), (7, 128)])
def test_layernorm_backward(M, D):
    g = torch.Generator().manual_seed(M)
    x = torch.randn(M, D, generator=g) * 2 + 0.3
    w = 1 + 0.1 * torch.randn(D, generator=g)
    b = 0.1 * torch.randn(D, generator=g)
    dy = torch.randn(M, D, generator=g)
    x0, w0, b0 = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    F.layer_norm(x0, (D,), w0, b0, 1e-5).backward(dy)
    dx, dg, db = _ops().layernorm_backward(x.cuda(), w.cuda(), dy.cuda(), 1e-5)
    _close(dx, x0.grad, "dx", 5e-5)
    _close(dg, w0.grad, "dgamma", 1e-4)
    _close(db, b0.grad, "dbeta", 1e-4)


@pytest.mark.parametrize("B,Q,heads", [(3, 13, 8), (2, 5, 2), (1, 16, 1), (4, 1, 2)])
def test_self_attention_backward(B, Q, heads):
    g = torch.Generator().manual_seed(B * 10 + Q)
    Cc = heads * 64
    qkv = torch.randn(B * Q, 3 * Cc, generator=g) * 0.5
    dO = torch.randn(B * Q, Cc, generator=g)
    t = qkv.clone().requires_grad_(True)
    x = t.view(B, Q, 3, heads, 64).permute(2, 0, 3, 1, 4)
    o = (torch.softmax(x[0] @ x[1].transpose(-1, -2), -1) @ x[2]).permute(0, 2, 1, 3).reshape(B * Q, Cc)
    o.backward(dO)
    got = _ops().self_attention_backward(qkv.cuda(), dO.cuda(), B, Q, heads)
    _close(got, t.grad, "dqkv", 5e-5)


@pytest.mark.parametrize("B,Q,heads,S", [(2, 13, 8, 1024), (1, 5, 2, 196), (3, 13, 2, 77), (1, 16, 1, 300)])
def test_cross_attention_backward(B, Q, heads, S):
    g = torch.Generator().manual_seed(S + Q)
    Cc = heads * 64
    q = torch.randn(B * Q, Cc, generator=g) * 0.3
    K = torch.randn(B * S, Cc, generator=g).to(torch.bfloat16)
    V = torch.randn(B * S, Cc, generator=g).to(torch.bfloat16)
    dO = torch.randn(B * Q, Cc, generator=g)
    q0, K0, V0 = q.clone().requires_grad_(True), K.float().requires_grad_(True), V.float().requires_grad_(True)
    qh = q0.view(B, Q, heads, 64).transpose(1, 2)
    kh, vh = K0.view(B, S, heads, 64).transpose(1, 2), V0.view(B, S, heads, 64).transpose(1, 2)
    O = (torch.softmax(qh @ kh.transpose(-1, -2), -1) @ vh).transpose(1, 2).reshape(B * Q, Cc)
    O.backward(dO)
    dq, dK, dV = _ops().cross_attention_backward(q.cuda(), K.cuda(), V.cuda(), O.detach().cuda(), dO.cuda(), B, Q, heads, S)
    _close(dq, q0.grad, "dq", 1e-4)
    _close(dK, K0.grad, "dK", 2 ** -7, atol=1e-4)
    _close(dV, V0.grad, "dV", 2 ** -7, atol=1e-4)

"""Tensor-level wrappers over the C ABI (one Python function per exported operator).

PyTorch is used for device memory and streams only; every function enqueues hand-written sm_100a kernels on the
current CUDA stream through libhh_b200.so.
"""
from __future__ import annotations

import torch

from . import _lib as L

EPI_BIAS_BF16, EPI_BIAS_QGELU_BF16, EPI_BIAS_RES_F32, EPI_BIAS_F32 = 0, 1, 2, 3


def _c(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous() else t.contiguous()


def _f32(t: torch.Tensor) -> torch.Tensor:
    return _c(t if t.dtype == torch.float32 else t.float())


# ------------------------------------------------------------------------------------------ kernel-level
def gemm_bf16(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None = None, epilogue: int = EPI_BIAS_BF16,
              residual: torch.Tensor | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
    """epilogue(a[M,K] @ w[N,K]^T): a, w bf16; bias fp32 [N]; residual fp32 [M,N] (may be `out`)."""
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.dim() == 2 and w.dim() == 2
    a, w = _c(a), _c(w)
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K
    odt = torch.bfloat16 if epilogue in (EPI_BIAS_BF16, EPI_BIAS_QGELU_BF16) else torch.float32
    if out is None:
        out = torch.empty(M, N, dtype=odt, device=a.device)
    assert out.dtype == odt and out.shape == (M, N) and out.is_contiguous()
    L.check(L.load().hh_gemm_bf16(L.ptr(a), K, L.ptr(w), K, L.ptr(out), N, L.ptr(bias), L.ptr(residual),
                                  N if residual is not None else 0, M, N, K, epilogue, L.stream_ptr()), "hh_gemm_bf16")
    return out


def fold_layernorm_weight(w: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, bias: torch.Tensor | None = None,
                          scaled_rows: int = 0, scale: float = 1.0):
    """LayerNorm(gamma, beta) folded into the Linear(w, bias) that follows it -> (Wf bf16 [N,K], colsum [N], bias_f [N])."""
    w32, g32, b32 = _f32(w), _f32(gamma), _f32(beta)
    bias32 = _f32(bias) if bias is not None else None
    N, K = w32.shape
    wf = torch.empty(N, K, dtype=torch.bfloat16, device=w.device)
    cs = torch.empty(N, dtype=torch.float32, device=w.device)
    bf = torch.empty(N, dtype=torch.float32, device=w.device)
    L.check(L.load().hh_fold_layernorm_weight(L.ptr(w32), L.ptr(g32), L.ptr(b32), L.ptr(bias32), N, K, scaled_rows, scale,
                                              L.ptr(wf), L.ptr(cs), L.ptr(bf), L.stream_ptr()), "hh_fold_layernorm_weight")
    return wf, cs, bf


def gemm_bf16_res_stats(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None, residual: torch.Tensor,
                        writeback: bool = False):
    """Producer side of the folded LayerNorm: z = a w^T + bias + residual -> (z bf16 [M,N], stats fp32 [parts, M, 2]);
    `residual` (fp32 [M,N]) receives z in place when `writeback`."""
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and residual.dtype == torch.float32
    a, w = _c(a), _c(w)
    assert residual.is_contiguous()
    M, K = a.shape
    N = w.shape[0]
    assert residual.shape == (M, N)
    parts = L.load().hh_gemm_stats_parts(M, N)
    z = torch.empty(M, N, dtype=torch.bfloat16, device=a.device)
    stats = torch.zeros(parts, M, 2, dtype=torch.float32, device=a.device)
    bias32 = _f32(bias) if bias is not None else None
    L.check(L.load().hh_gemm_bf16_res_stats(L.ptr(a), K, L.ptr(w), K, L.ptr(z), N, L.ptr(bias32), L.ptr(residual), N,
                                            1 if writeback else 0, L.ptr(stats), M, N, K, L.stream_ptr()),
            "hh_gemm_bf16_res_stats")
    return z, stats


def gemm_bf16_ln(z: torch.Tensor, wf: torch.Tensor, bias_f: torch.Tensor, colsum: torch.Tensor, stats: torch.Tensor,
                 eps: float, qgelu: bool = False) -> torch.Tensor:
    """Consumer side: act(Linear(LayerNorm(z))) with the norm folded into (wf, colsum, bias_f); stats [parts, M, 2]."""
    assert z.dtype == torch.bfloat16 and wf.dtype == torch.bfloat16 and stats.dtype == torch.float32
    z, wf, stats = _c(z), _c(wf), _c(stats)
    M, K = z.shape
    N = wf.shape[0]
    assert stats.dim() == 3 and stats.shape[1] == M and stats.shape[2] == 2
    out = torch.empty(M, N, dtype=torch.bfloat16, device=z.device)
    b32, c32 = _f32(bias_f), _f32(colsum)
    L.check(L.load().hh_gemm_bf16_ln(L.ptr(z), K, L.ptr(wf), K, L.ptr(out), N, L.ptr(b32), L.ptr(c32), L.ptr(stats),
                                     stats.shape[0], K, eps, M, N, K, 1 if qgelu else 0, L.stream_ptr()), "hh_gemm_bf16_ln")
    return out


def layernorm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float, want_f32=True, want_bf16=False):
    x = _f32(x)
    M, D = x.shape
    o32 = torch.empty_like(x) if want_f32 else None
    o16 = torch.empty(M, D, dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    w32, b32 = _f32(w), _f32(b)     # bound to locals: a converted temporary must outlive the call that reads its pointer
    L.check(L.load().hh_layernorm(L.ptr(x), D, L.ptr(w32), L.ptr(b32), eps, L.ptr(o32), L.ptr(o16), M, D,
                                  L.stream_ptr()), "hh_layernorm")
    return o32, o16


def attention(qkv: torch.Tensor, B: int, T: int, n: int, H: int, standalone_cls: bool = False) -> dict:
    """Divided space-time attention on packed qkv (bf16 [B*(1+T*n), 3*H*64], q pre-scaled).
    Returns {'space': o, 'time': o} with every row filled (patch rows + the CLS query row).  With
    `standalone_cls` the CLS row is overwritten by the stand-alone CLS kernel (kind 2)."""
    assert qkv.dtype == torch.bfloat16 and qkv.is_contiguous()
    N = 1 + T * n
    outs = {}
    lib = L.load()
    for kind, name in ((0, "space"), (1, "time")):
        o = torch.zeros(B * N, H * 64, dtype=torch.bfloat16, device=qkv.device)
        L.check(lib.hh_attention(L.ptr(qkv), L.ptr(o), B, T, n, H, kind, L.stream_ptr()), "hh_attention")
        if standalone_cls:
            L.check(lib.hh_attention(L.ptr(qkv), L.ptr(o), B, T, n, H, 2, L.stream_ptr()), "hh_attention(cls)")
        outs[name] = o
    return outs


def attention_causal(qkv: torch.Tensor, G: int, Lc: int, H: int) -> torch.Tensor:
    """Causal self-attention of the text tower on packed qkv (bf16 [G*Lc, 3*H*64], q pre-scaled) -> bf16 [G*Lc, H*64]."""
    assert qkv.dtype == torch.bfloat16 and qkv.is_contiguous() and qkv.shape == (G * Lc, 3 * H * 64)
    o = torch.empty(G * Lc, H * 64, dtype=torch.bfloat16, device=qkv.device)
    L.check(L.load().hh_attention_causal(L.ptr(qkv), L.ptr(o), G, Lc, H, L.stream_ptr()), "hh_attention_causal")
    return o


def cross_attention(q: torch.Tensor, K: torch.Tensor, V: torch.Tensor, B: int, Q: int, heads: int, S: int,
                    simt: bool = False) -> torch.Tensor:
    """q fp32 [B*Q, heads*64] (pre-scaled); K, V bf16 [B*S, heads*64].  `simt` selects the fp32 SIMT statement."""
    q = _f32(q)
    assert K.dtype == torch.bfloat16 and V.dtype == torch.bfloat16 and K.is_contiguous() and V.is_contiguous()
    out = torch.empty_like(q)
    fn = L.load().hh_cross_attention_simt if simt else L.load().hh_cross_attention
    L.check(fn(L.ptr(q), L.ptr(K), L.ptr(V), K.shape[1], L.ptr(out), B, Q, heads, S, L.stream_ptr()),
            "hh_cross_attention")
    return out


# ------------------------------------------------------------------------------------------ operators
def linear_f32(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None = None, act: int = 0,
               in_relu: bool = False) -> torch.Tensor:
    """nn.Linear on the last dim (fp32), optional input ReLU and output ReLU(1)/sigmoid(2)."""
    shp = x.shape
    x2 = _f32(x).reshape(-1, shp[-1])
    N, K = weight.shape
    out = torch.empty(x2.shape[0], N, dtype=torch.float32, device=x.device)
    w32 = _f32(weight)              # locals keep converted (non-fp32 / non-contiguous) parameters alive across the call
    b32 = _f32(bias) if bias is not None else None
    if x2.shape[0] > 0:
        L.check(L.load().hh_linear_f32(L.ptr(x2), K, None, 0, L.ptr(w32), L.ptr(b32),
                                       None, 0, L.ptr(out), N, x2.shape[0], N, K, act, 1 if in_relu else 0,
                                       L.stream_ptr()), "hh_linear_f32")
    return out.reshape(*shp[:-1], N)


def l2_normalize(x: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    shp = x.shape
    x2 = _f32(x).reshape(-1, shp[-1])
    out = torch.empty_like(x2)
    L.check(L.load().hh_l2_normalize(L.ptr(x2), L.ptr(out), x2.shape[0], x2.shape[1], eps, L.stream_ptr()),
            "hh_l2_normalize")
    return out.reshape(shp)


def sim_matrix(a: torch.Tensor, b: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    a, b = _f32(a), _f32(b)
    Na, d = a.shape
    Nb = b.shape[0]
    out = torch.empty(Na, Nb, dtype=torch.float32, device=a.device)
    L.check(L.load().hh_sim_matrix(L.ptr(a), L.ptr(b), L.ptr(out), Na, Nb, d, eps, L.stream_ptr()), "hh_sim_matrix")
    return out


def row_argmax(x: torch.Tensor) -> torch.Tensor:
    x = _f32(x)
    out = torch.empty(x.shape[0], dtype=torch.int64, device=x.device)
    L.check(L.load().hh_row_reduce(L.ptr(x), x.shape[0], x.shape[1], 1.0, 0, L.ptr(out), L.stream_ptr()), "hh_row_reduce")
    return out


def row_softmax(x: torch.Tensor, scale: float = 1.0, log: bool = False) -> torch.Tensor:
    x = _f32(x)
    out = torch.empty_like(x)
    L.check(L.load().hh_row_reduce(L.ptr(x), x.shape[0], x.shape[1], scale, 2 if log else 1, L.ptr(out),
                                   L.stream_ptr()), "hh_row_reduce")
    return out


def box_convert(x: torch.Tensor, to_xyxy: bool) -> torch.Tensor:
    shp = x.shape
    assert shp[-1] == 4
    x2 = _f32(x).reshape(-1, 4)
    out = torch.empty_like(x2)
    fn = L.load().hh_box_cxcywh_to_xyxy if to_xyxy else L.load().hh_box_xyxy_to_cxcywh
    if x2.shape[0] > 0:
        L.check(fn(L.ptr(x2), L.ptr(out), x2.shape[0], L.stream_ptr()), "hh_box_convert")
    return out.reshape(shp)


def box_pairwise(b1: torch.Tensor, b2: torch.Tensor):
    """(iou, union, giou), each [N,M], for xyxy boxes."""
    b1, b2 = _f32(b1), _f32(b2)
    N, M = b1.shape[0], b2.shape[0]
    iou = torch.empty(N, M, dtype=torch.float32, device=b1.device)
    uni = torch.empty_like(iou)
    giou = torch.empty_like(iou)
    if N > 0 and M > 0:
        L.check(L.load().hh_box_pairwise(L.ptr(b1), L.ptr(b2), N, M, L.ptr(iou), L.ptr(uni), L.ptr(giou), L.stream_ptr()),
                "hh_box_pairwise")
    return iou, uni, giou


def box_match_cost(pred: torch.Tensor, tgt: torch.Tensor, w_bbox: float = 5.0, w_giou: float = 2.0) -> torch.Tensor:
    pred, tgt = _f32(pred), _f32(tgt)
    N, M = pred.shape[0], tgt.shape[0]
    cost = torch.empty(N, M, dtype=torch.float32, device=pred.device)
    if N > 0 and M > 0:
        L.check(L.load().hh_box_match_cost(L.ptr(pred), L.ptr(tgt), N, M, w_bbox, w_giou, L.ptr(cost), L.stream_ptr()),
                "hh_box_match_cost")
    return cost


def match_cost_class(cost: torch.Tensor, logits: torch.Tensor, tgt_ids: torch.Tensor, weight: float) -> torch.Tensor:
    """cost[r, t] += weight * -softmax(logits[r])[tgt_ids[t]] in place (matcher class term, reference
    model/box_utils.py:66,83-85)."""
    logits = _f32(logits)
    ids = tgt_ids.to(torch.int64).contiguous()
    N, M = cost.shape
    assert logits.shape[0] == N and ids.numel() == M and cost.dtype == torch.float32 and cost.is_contiguous()
    if N > 0 and M > 0:
        L.check(L.load().hh_match_cost_class(L.ptr(logits), N, logits.shape[1], L.ptr(ids), M, float(weight), L.ptr(cost),
                                             L.stream_ptr()), "hh_match_cost_class")
    return cost


def assign(cost: torch.Tensor, offset, ld, nr, nc, row_valid: torch.Tensor | None = None):
    """Batched exact linear-sum assignment on the device (scipy.optimize.linear_sum_assignment semantics / index order).
    Problem p = the nr[p] x nc[p] block of the fp32 tensor `cost` starting at element offset[p] with row stride ld[p];
    `row_valid` (uint8/bool [P, max nr]) drops rows first.  offset/ld/nr/nc are host sequences.
    Returns (row_ind int64 [P, K], col_ind int64 [P, K], count int32 [P]) on the device, -1 padded."""
    P = len(nr)
    assert cost.dtype == torch.float32 and cost.is_contiguous() and cost.is_cuda
    max_dim = max([1] + [int(v) for v in nr] + [int(v) for v in nc])
    if max_dim > 32:
        raise NotImplementedError("assignment problems larger than 32 x 32 are not supported (got %d)" % max_dim)
    K = max(1, min(max([0] + [int(v) for v in nr]), max([0] + [int(v) for v in nc])))
    dev = cost.device
    ri = torch.full((P, K), -1, dtype=torch.int64, device=dev)
    ci = torch.full((P, K), -1, dtype=torch.int64, device=dev)
    cnt = torch.zeros(P, dtype=torch.int32, device=dev)
    if P == 0:
        return ri, ci, cnt
    meta64 = torch.tensor([int(v) for v in offset], dtype=torch.int64).to(dev, non_blocking=True)
    meta32 = torch.tensor([[int(v) for v in ld], [int(v) for v in nr], [int(v) for v in nc]],
                          dtype=torch.int32).to(dev, non_blocking=True)
    rv, rv_ld = None, 0
    if row_valid is not None:
        rv = row_valid.to(torch.uint8).contiguous()
        assert rv.shape[0] == P
        rv_ld = rv.shape[1]
    L.check(L.load().hh_assign(L.ptr(cost), L.ptr(meta64), L.ptr(meta32[0]), L.ptr(meta32[1]), L.ptr(meta32[2]),
                               L.ptr(rv), rv_ld, P, max_dim, L.ptr(ri), L.ptr(ci), L.ptr(cnt), K, L.stream_ptr()),
            "hh_assign")
    return ri, ci, cnt


def sim_matrix_backward(a: torch.Tensor, b: torch.Tensor, grad: torch.Tensor, eps: float = 1e-8, need_a: bool = True,
                        need_b: bool = True):
    """Gradient of sim_matrix(a, b) w.r.t. a and b (None where not needed)."""
    a, b, grad = _f32(a), _f32(b), _f32(grad)
    Na, d = a.shape
    Nb = b.shape[0]
    assert grad.shape == (Na, Nb)
    lib = L.load()
    ws = torch.empty(lib.hh_sim_matrix_backward_workspace_bytes(Na, Nb), dtype=torch.uint8, device=a.device)
    da = torch.empty_like(a) if need_a else None
    db = torch.empty_like(b) if need_b else None
    L.check(lib.hh_sim_matrix_backward(L.ptr(a), L.ptr(b), L.ptr(grad), None, L.ptr(da), L.ptr(db), Na, Nb, d, eps,
                                       L.ptr(ws), L.stream_ptr()), "hh_sim_matrix_backward")
    return da, db


def retrieval_rows(sim, rel, mode: int, kcounts=None):
    """Per-query average precision (mode 0, utils/mAP.py) or DCG (mode 1, utils/nDCG.py) on the device, float64.
    `sim`, `rel` [N, M]: numpy arrays or tensors (moved to the current CUDA device).  Returns a float64 numpy vector [N]."""
    import numpy as np

    def dev64(x):
        t = torch.as_tensor(np.ascontiguousarray(x)) if isinstance(x, np.ndarray) else x
        return t.to(device="cuda", dtype=torch.float64).contiguous()
    s, r = dev64(sim), dev64(rel)
    assert s.dim() == 2 and s.shape == r.shape, "similarity and relevancy matrices must have the same 2-D shape"
    N, M = s.shape
    out = torch.empty(N, dtype=torch.float64, device=s.device)
    logs = kc = None
    if mode == 1:
        logs = torch.from_numpy(np.log2(np.arange(M) + 2)).to(s.device)          # numpy's own divisor table
        if kcounts is not None:
            kc = torch.as_tensor(np.ascontiguousarray(kcounts) if isinstance(kcounts, np.ndarray) else kcounts)
            kc = kc.to(device=s.device, dtype=torch.int32).contiguous()
            assert kc.shape == s.shape
    L.check(L.load().hh_retrieval_rows(L.ptr(s), L.ptr(r), L.ptr(logs), L.ptr(kc), N, M, mode, L.ptr(out),
                                       L.stream_ptr()), "hh_retrieval_rows")
    return out.cpu().numpy()


# ------------------------------------------------------------------------------------------ backward primitives
def linear_f32_backward(dy, y, act, weight, x, x_add=None, in_relu=False, need_dx=True, need_dw=True):
    """Backward of linear_f32: returns (dx, dw, db)."""
    dy, weight, x = _f32(dy), _f32(weight), _f32(x)
    R, N = dy.shape
    K = weight.shape[1]
    dx = torch.empty(R, K, dtype=torch.float32, device=dy.device) if need_dx else None
    dw = torch.empty(N, K, dtype=torch.float32, device=dy.device) if need_dw else None
    db = torch.empty(N, dtype=torch.float32, device=dy.device) if need_dw else None
    yy = _f32(y) if act else None
    xa = _f32(x_add) if x_add is not None else None
    L.check(L.load().hh_linear_f32_backward(L.ptr(dy), N, L.ptr(yy), N, act, L.ptr(weight), L.ptr(x), K, L.ptr(xa),
                                            xa.shape[0] if xa is not None else 0, 1 if in_relu else 0, L.ptr(dx), K,
                                            L.ptr(dw), L.ptr(db), R, N, K, 0.0, 1.0, L.stream_ptr()),
            "hh_linear_f32_backward")
    return dx, dw, db


def layernorm_backward(x, w, dy, eps):
    x, w, dy = _f32(x), _f32(w), _f32(dy)
    M, D = x.shape
    dx, dg, db = torch.empty_like(x), torch.empty_like(w), torch.empty_like(w)
    L.check(L.load().hh_layernorm_backward(L.ptr(x), D, L.ptr(w), eps, L.ptr(dy), D, L.ptr(dx), L.ptr(dg), L.ptr(db), M, D,
                                           L.stream_ptr()), "hh_layernorm_backward")
    return dx, dg, db


def self_attention_backward(qkv, dO, B, Q, heads):
    """qkv fp32 [B*Q, 3C] packed (q pre-scaled) -> d(qkv) in the same packing."""
    qkv, dO = _f32(qkv), _f32(dO)
    Cc = heads * 64
    g = torch.empty_like(qkv)
    p, pg = L.ptr(qkv), L.ptr(g)
    L.check(L.load().hh_self_attention_backward(p, p + 4 * Cc, p + 8 * Cc, 3 * Cc, L.ptr(dO), pg, pg + 4 * Cc, pg + 8 * Cc,
                                                3 * Cc, B, Q, heads, L.stream_ptr()), "hh_self_attention_backward")
    return g


def cross_attention_backward(q, K, V, O, dO, B, Q, heads, S):
    q, O, dO = _f32(q), _f32(O), _f32(dO)
    dq = torch.empty_like(q)
    dK, dV = torch.empty_like(K), torch.empty_like(V)
    L.check(L.load().hh_cross_attention_backward(L.ptr(q), L.ptr(K), L.ptr(V), K.shape[1], L.ptr(O), L.ptr(dO), L.ptr(dq),
                                                 L.ptr(dK), L.ptr(dV), dK.shape[1], B, Q, heads, S, L.stream_ptr()),
            "hh_cross_attention_backward")
    return dq, dK, dV
